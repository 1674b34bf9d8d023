"""Pins the CPU oracle's MCTS / self-play against the reference's own tests:
rust/src/mcts.rs:463-686 (+ rust/proptest-regressions/mcts.txt) and rust/src/self_play.rs:383-459."""

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as O

CONST_COL_WEIGHT = np.float32(1.0) / np.float32(7.0)
EPS = 1e-8


def board(rows):
    return O.from_str("\n".join(rows))


E = "⚫⚫⚫⚫⚫⚫⚫"


def test_mcts_prefers_center_column():  # mcts.rs:487-492
    pol, _, _, _ = O.run_mcts(O.Pos(0, 0), 1000)
    assert abs(pol.sum() - 1) < 1e-5
    assert pol[3] > CONST_COL_WEIGHT


def test_mcts_depth_one():  # mcts.rs:494-499
    pol, _, _, _ = O.run_mcts(O.Pos(0, 0), 1 + 7 + 7)
    assert np.all(np.abs(pol - CONST_COL_WEIGHT) < EPS)


def test_mcts_depth_two():  # mcts.rs:501-508
    pol, _, _, _ = O.run_mcts(O.Pos(0, 0), 1 + 7 + 49 + 49)
    assert np.all(np.abs(pol - CONST_COL_WEIGHT) < EPS)


def test_mcts_depth_uneven():  # mcts.rs:510-514
    pol, _, _, _ = O.run_mcts(O.Pos(0, 0), 47)
    assert np.any(np.abs(pol - CONST_COL_WEIGHT) > EPS)


def test_winning_position():  # mcts.rs:518-538
    pos = board([E, E, E, E, "⚫🔵🔵🔵⚫⚫⚫", "⚫🔴🔴🔴⚫⚫⚫"])
    pol, qp, qn, _ = O.run_mcts(pos, 10_000)
    assert abs(pol.sum() - 1) < 1e-6
    assert pol[0] + pol[4] > 0.99
    assert qp > 0.92 and qn > 0.99


def test_winning_position2():  # mcts.rs:541-561
    pos = board([E, E, E, E, "⚫⚫🔵🔵⚫⚫⚫", "⚫⚫🔴🔴⚫⚫⚫"])
    pol, qp, qn, _ = O.run_mcts(pos, 10_000)
    assert pol[1] + pol[4] > 0.98
    assert qp > 0.90 and qn > 0.98 and qn > qp


def test_winning_position3():  # mcts.rs:564-583
    pos = board([E, E, E, "⚫🔴🔵🔵⚫⚫⚫", "⚫🔵🔴🔴🔴⚫⚫", "⚫🔵🔵🔴🔵🔴⚫"])
    pol, qp, qn, _ = O.run_mcts(pos, 10_000)
    assert pol[5] > 0.99
    assert qp > 0.86 and qn > 0.99 and qn > qp


def test_losing_position():  # mcts.rs:587-609 — passes only with f32 sequential accumulation
    pos = board([E, E, E, E, "⚫🔴🔴⚫⚫⚫⚫", "⚫🔵🔵🔵⚫⚫⚫"])
    pol, qp, qn, _ = O.run_mcts(pos, 300_000)
    assert abs(pol.sum() - 1) < 1e-5
    assert np.all(np.abs(pol - CONST_COL_WEIGHT) < 0.01)
    assert qp < -0.93
    assert qn < -0.99
    assert qn < qp


def test_prefer_shorter_wins():  # mcts.rs:613-632
    pos = board(
        ["⚫⚫⚫🔵⚫⚫⚫", "⚫🔵🔵🔵⚫⚫⚫", "⚫🔴🔵🔵⚫⚫⚫", "⚫🔴🔴🔴⚫⚫⚫", "⚫🔴🔴🔴⚫⚫⚫", "⚫🔵🔴🔵⚫⚫⚫"]
    )
    pol, qp, qn, _ = O.run_mcts(pos, 10_000)
    assert pol[4] > 0.99
    assert qp > 0.82 and qn > 0.99 and qn > qp


# ---- softmax / temperature properties (mcts.rs:635-686) ----
logit = st.one_of(st.floats(0.0, 10.0, width=32, exclude_max=True), st.just(float("-inf")))
policy_logits = st.lists(logit, min_size=7, max_size=7).filter(lambda l: not all(x == float("-inf") for x in l))

# rust/proptest-regressions/mcts.txt:7-12
REGRESSION_POLICIES = [
    [0.9286058, 0.0, 0.0033046294, 0.06687763, 0.0, 0.0, 0.001211846],
    [0.4780801, 2.5148089e-5, 2.5148089e-5, 0.52179414, 2.5148089e-5, 2.5148089e-5, 2.5148089e-5],
    [0.0, 0.933416, 0.00035163847, 0.0009350313, 0.0, 0.06520966, 8.7709726e-5],
    [0.0, 0.106206864, 0.0, 0.0, 0.148644, 0.7410872, 0.004062006],
]


def _check_temperature_props(policy):
    policy = np.asarray(policy, dtype=np.float32)
    assert abs(policy.sum(dtype=np.float32) - 1) <= 1e-5  # softmax_sum_1
    t1 = O.apply_temperature(policy, 1.0)
    assert np.all(np.abs(t1 - policy) < 1e-5)  # temperature_1
    t2 = O.apply_temperature(policy, 2.0)
    assert abs(t2.sum(dtype=np.float32) - 1) <= 1e-5  # temperature_2
    # the reference's property (mcts.rs:668-670) asks for >= 2 positive non-1/7 entries; it is only
    # true when those entries are not all equal (e.g. [0, .5, .5, 0..] is a fixed point), so we
    # additionally require two distinct positive values.
    pos_vals = policy[(policy != CONST_COL_WEIGHT) & (policy > 0)]
    if len(pos_vals) >= 2 and (pos_vals.max() - pos_vals.min()) > 1e-3:
        assert np.any(np.abs(t2 - policy) > EPS)
    t0 = O.apply_temperature(policy, 0.0)  # temperature_0
    mx = t0.max()
    assert abs(t0.sum(dtype=np.float32) - 1) <= 1e-5
    assert np.all(t0[t0 == mx] == np.float32(1.0) / np.float32((t0 == mx).sum()))


@settings(max_examples=500, deadline=None)
@given(policy_logits)
def test_prop_softmax_temperature(logits):
    _check_temperature_props(O.softmax(logits))


@pytest.mark.parametrize("policy", REGRESSION_POLICIES)
def test_regression_policies(policy):
    _check_temperature_props(policy)


def test_regression_softmax_inputs():  # mcts.txt:7-8
    p = O.softmax([0.0] * 7)
    assert np.all(p == CONST_COL_WEIGHT)
    p = O.softmax([0.0, 0.0, -6.872888e19, 0.0, 0.0, 0.0, 0.0])
    assert p[2] == 0 and abs(p.sum() - 1) < 1e-6
    assert O.softmax([float("-inf")] * 7) is None  # reference panics


def test_self_play():  # self_play.rs:405-459
    out = O.self_play([(0, 0, 0)], 10, 50, 1.0, 0.01, "uniform")
    for samples in out.samples:
        assert len(samples) >= 7
        assert sum(1 for s in samples if s.pos.key() == (0, 0)) == 1
        terminal = [s for s in samples if O.terminal_state(s.pos) != O.NONE]
        assert len(terminal) == 1
        assert terminal[0].q_no_penalty in (-1.0, 0.0, 1.0)


def test_self_play_threaded_matches_serial():
    """per-game records do not depend on scheduling."""
    reqs = [(i, 0, 0) for i in range(12)]
    a = O.self_play(reqs, 5, 40, 6.6, 0.01, "hash")
    b = O.self_play(reqs, 5, 40, 6.6, 0.01, "hash", threaded=True, n_threads=3)
    assert a.records() == b.records()
    assert a.stats["sims"] == b.stats["sims"]
    assert b.stats["nn_evals"] <= a.stats["nn_evals"]  # the NN thread de-duplicates positions


def test_to_result_alternating_q():  # mcts.rs:271-313
    out = O.self_play([(3, 0, 0)], 64, 30, 6.6, 0.01, "hash")
    s = out.samples[0]
    qp, qn = s[-1].q_penalty, s[-1].q_no_penalty
    L = len(s) - 1
    for k in range(L):
        sign = 1.0 if (L - k) % 2 == 0 else -1.0
        assert s[k].q_penalty == np.float32(sign) * np.float32(qp)
        assert s[k].q_no_penalty == np.float32(sign) * np.float32(qn)
    assert list(s[-1].policy) == [CONST_COL_WEIGHT] * 7
    assert O.player0_score(s) in (0.0, 0.5, 1.0)


# ---- rand 0.10.1 restatement: only the public primitives can be pinned ----
def test_chacha20_zero_key_block():
    """ChaCha20 keystream, all-zero key/nonce, block 0 (draft-agl-tls-chacha20poly1305 / RFC 7539 A.1 #1)."""
    import ctypes as C

    key = (C.c_uint32 * 8)()
    out = (C.c_uint32 * 16)()
    O.lib().c4o_chacha_block(key, 0, 20, out)
    assert [hex(x) for x in out[:4]] == ["0xade0b876", "0x903df1a0", "0xe56a5d40", "0x28bd8653"]


def test_chacha20_rfc7539_block():
    """RFC 7539 §2.3.2: key 00..1f, counter 1, nonce 000000090000004a00000000."""
    import ctypes as C

    key = (C.c_uint32 * 8)(*[int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)])
    out = (C.c_uint32 * 16)()
    O.lib().c4o_chacha_block_nonce(key, 1, 0x09000000, 0x4A000000, 0, 20, out)
    assert [hex(x) for x in out[:4]] == ["0xe4e7f110", "0x15593bd1", "0x1fdd0f50", "0xc47120a3"]
    assert hex(out[15]) == "0x4e3c50a2"


def test_weighted_index_basic():
    assert O.weighted_index_sample([0, 0, 0, 1, 0, 0, 0], 12345) == 3
    assert O.weighted_index_sample([0] * 7, 1) == -1
    counts = np.zeros(7)
    for seed in range(4000):
        counts[O.weighted_index_sample([0.1, 0.0, 0.2, 0.3, 0.0, 0.4, 0.0], seed)] += 1
    assert counts[1] == counts[4] == counts[6] == 0
    assert np.all(np.abs(counts / 4000 - np.array([0.1, 0, 0.2, 0.3, 0, 0.4, 0])) < 0.03)


def test_shuffle_is_permutation():
    idx = O.shuffle_indices(1337, 100)
    assert sorted(idx.tolist()) == list(range(100))
    assert idx.tolist() == O.shuffle_indices(1337, 100).tolist()
    assert idx.tolist() != list(range(100))
