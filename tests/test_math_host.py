"""Host builds of the product's shared math (same source as the device code) against libm and the
oracle.  The exhaustive 2^32 sweep of c4_logf / c4_expf vs libm is recorded in DESIGN.md §4; here a
few million inputs keep the CPU suite fast."""

import numpy as np
import pytest

import oracle
from c4a0_b200 import engine as E


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.mark.parametrize("seed", [0, 1])
def test_logf_expf_bit_exact_vs_libm(seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2**32, size=3_000_000, dtype=np.uint64).astype(np.uint32)
    x = np.concatenate([
        bits.view(np.float32),
        np.arange(0, 200_000, dtype=np.float32),                       # ln(visit counts)
        (-rng.random(1_000_000) * 110).astype(np.float32),             # softmax arguments
        (rng.random(500_000)).astype(np.float32),                      # ln(policy)
        np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, 1e-45, 88.7, 88.8, -103.9, -104.0, -87.3], np.float32),
    ])
    for mine, ref in ((E.host_logf, oracle.logf), (E.host_expf, oracle.expf)):
        got, exp = mine(x), ref(x)
        ok = _same(got, exp)
        assert ok.all(), f"{np.count_nonzero(~ok)} mismatches, first x={x[~ok][0]!r}"


def test_temperature_and_sampling_match_oracle():
    rng = np.random.default_rng(5)
    for _ in range(4000):
        counts = rng.integers(0, 600, size=7).astype(np.float32)
        counts[rng.random(7) < 0.25] = 0
        if counts.sum() == 0:
            counts[int(rng.integers(0, 7))] = 3
        pol = (counts / np.float32(counts.sum())).astype(np.float32)
        t = float(rng.choice([4.0, 2.0, 1.0, 0.0]))
        seed = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        tempered, col = E.host_sample(pol, t, seed)
        exp_t = oracle.apply_temperature(pol.tolist(), t)
        assert tempered.tobytes() == exp_t.tobytes()
        assert col == oracle.weighted_index_sample(exp_t.tolist(), seed)
    # the reference panics on an all-zero / negative weight vector: we return -1
    assert E.host_sample(np.zeros(7, np.float32), 1.0, 3)[1] == -1
    assert E.host_sample(np.array([1, -1, 0, 0, 0, 0, 0], np.float32), 1.0, 3)[1] == -1


def test_game_id_zero_always_uses_seed_zero():
    # mcts.rs:215: seed = game_id * (42 + n_moves) — game 0 draws from seed 0 at every move
    w = np.full(7, 1 / 7, np.float32)
    assert len({E.host_sample(w, 1.0, 0 * (42 + k))[1] for k in range(10)}) == 1


def test_shuffle_matches_oracle():
    for seed in (0, 1, 1337, 2**64 - 1):
        for n in (0, 1, 2, 3, 12, 13, 100, 1700, 5000):
            assert np.array_equal(E.host_shuffle(seed, n), oracle.shuffle_indices(seed, n)), (seed, n)


def test_rules_host_functions_match_oracle():
    rng = np.random.default_rng(9)
    L = oracle.lib()
    for _ in range(3000):
        p = oracle.random_pos(rng.integers(0, 7, size=int(rng.integers(0, 60))).tolist())
        assert E.host_terminal_state(p.mask, p.value) == oracle.terminal_state(p)
        f = L.c4o_flip_h(p)
        assert E.host_flip_h(p.mask, p.value) == (f.mask, f.value)
        for c in range(7):
            ch = oracle.make_move(p, c)
            if ch is not None:
                import ctypes as C

                a, b = C.c_uint64(), C.c_uint64()
                from c4a0_b200 import _lib

                _lib.lib().c4a0_host_make_move(p.mask, p.value, c, C.byref(a), C.byref(b))
                assert (a.value, b.value) == ch.key()


def test_shift_and_win_test_equals_the_69_masks():
    """has_four (four shift-and tests) == the reference's WIN_MASKS scan, on arbitrary stone sets
    (not only reachable ones)."""
    rng = np.random.default_rng(11)
    masks = oracle.win_masks()
    for _ in range(20000):
        t = int(rng.integers(0, 2**42))
        if rng.random() < 0.5:
            t &= int(rng.integers(0, 2**42))
        want = any((t & m) == m for m in masks)
        # Player stones = t, no opponent stones: PLAYER_WIN iff a four exists (else NONE or DRAW)
        got = E.host_terminal_state(t, t) == 1
        assert got == want, hex(t)


def test_eval_cache_key_is_injective_and_decodes():
    """c4_rules.cuh pos_key(): 49 bits per position; checked on every position of 3,000 random games
    (reachable positions only: the key relies on gravity) by decoding it back to (mask, value)."""
    import oracle
    from c4a0_b200 import _lib as L

    lib = L.lib()
    rng = np.random.default_rng(11)
    seen = {}
    for _ in range(3000):
        p = oracle.Pos(0, 0)
        while True:
            key = int(lib.c4a0_host_pos_key(p.mask, p.value))
            assert key < (1 << 49)
            # decode: per column the highest set bit is the marker, below it 1 = side to move
            mask = value = 0
            for col in range(7):
                rows = [r for r in range(7) if (key >> (r * 7 + col)) & 1]
                top = max(rows)  # the marker always exists
                for r in range(top):
                    mask |= 1 << (r * 7 + col)
                    if (key >> (r * 7 + col)) & 1:
                        value |= 1 << (r * 7 + col)
            assert (mask, value) == p.key()
            assert seen.setdefault(key, p.key()) == p.key()
            if oracle.terminal_state(p) != 0:
                break
            legal = [c for c, ok in enumerate(oracle.legal_moves(p)) if ok]
            p = oracle.make_move(p, int(rng.choice(legal)))
    assert len(seen) > 20000
