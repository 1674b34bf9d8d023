"""Parity at the configurations the benchmark numbers are quoted on.

Every test here drives the SHIPPED path: `c4a0_rust.play_games` -> SelfPlaySession -> the native host
loop `c4a0_engine_run` over captured CUDA graphs, evaluation cache + speculative rows on, default arenas,
and compares complete game records per `game_id`, bit for bit, with the CPU oracle:

  * config 2: 16,384 lockstep games x 600 sims/move, the first 1,024 game ids in full;
  * tree state (structure and every N / Qp / Qn / prior bit pattern) of 64 games at n = 600 along the
    whole game, with the cache and speculation on;
  * 256 games at n = 1,400 (the reference CLI default, src/c4a0/main.py:41) and n = 1,600 (config 4).

The evaluators are the engine's integer-hash pseudo networks (tier E1): pure functions of
(model, position) that the oracle evaluates identically on the CPU.
"""

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

C_EXPL, C_PEN = 6.6, 0.01  # src/c4a0/main.py:42-43


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _records(soa, i):
    n = int(soa.n_samples[i])
    return [
        (
            int(soa.mask[i, k]),
            int(soa.value[i, k]),
            tuple(soa.policy[i, k].view(np.uint32).tolist()),
            int(soa.q_penalty[i, k].view(np.uint32)),
            int(soa.q_no_penalty[i, k].view(np.uint32)),
        )
        for k in range(n)
    ]


def _play(n_games, n_iter, kind, max_batch=None):
    import c4a0_rust
    from c4a0_b200.selfplay import BuiltinEvaluator

    reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in range(n_games)]
    # room for the speculative rows inside the caller's batch bound (see play_games)
    res = c4a0_rust.play_games(reqs, max_batch or (n_games + 8192), n_iter, C_EXPL, C_PEN, BuiltinEvaluator(kind))
    c4a0_rust._native.close_cached_session()
    return res


def test_config2_16384_games_600_sims_first_1024_ids_bit_exact():
    """BASELINE configs[1] on the shipped path; the oracle replays game ids 0..1023."""
    _need_gpu()
    n_games, n_iter, n_check = 16384, 600, 1024
    res = _play(n_games, n_iter, "hash_flat")
    info = res._run_info
    st = info.stats
    assert int((res._soa.n_samples > 0).sum()) == n_games
    assert st["cache_hits"] > 0 and st["spec_rows"] > 0, "the cache and speculation must be on for this test"
    assert info.report["ticks"] > 0  # native loop
    exp = oracle.self_play_parallel([(i, 0, 0) for i in range(n_check)], n_iter, C_EXPL, C_PEN, "hash_flat")
    for i in range(n_check):
        assert _records(res._soa, i) == exp[i], f"game_id {i}"
    # size-independent properties over all 16,384 games: every game ends in a terminal position whose
    # value is the objective result, samples alternate sign back to the start (mcts.rs:271-313)
    soa = res._soa
    last = soa.n_samples.astype(np.int64) - 1
    idx = np.arange(n_games)
    term = np.array([oracle.terminal_state(oracle.Pos(int(m), int(v))) for m, v in
                     zip(soa.mask[idx[::16], last[::16]], soa.value[idx[::16], last[::16]])])
    assert (term != oracle.NONE).all()
    qn_last = soa.q_no_penalty[idx, last]
    assert np.isin(qn_last, (-1.0, 0.0, 1.0)).all()
    first_qn = soa.q_no_penalty[:, 0]
    flip = np.where(last % 2 == 0, 1.0, -1.0).astype(np.float32)
    assert np.array_equal(first_qn, qn_last * flip)
    assert (soa.mask[:, 0] == 0).all()  # every game starts from the empty board


@pytest.mark.parametrize("n_iter", [1400, 1600])
def test_256_games_at_cli_default_and_config4_sims(n_iter):
    _need_gpu()
    n_games = 256
    res = _play(n_games, n_iter, "hash")
    assert res._run_info.stats["cache_hits"] > 0
    exp = oracle.self_play_parallel([(i, 0, 0) for i in range(n_games)], n_iter, C_EXPL, C_PEN, "hash", chunk=8)
    for i in range(n_games):
        assert _records(res._soa, i) == exp[i], f"game_id {i}"


def test_reference_cli_batch_1700_games_1400_sims_subset():
    """`main.py train` defaults: 1,700 games, 1,400 sims, batch 2,000 (src/c4a0/main.py:40-45)."""
    _need_gpu()
    n_games, n_iter, n_check = 1700, 1400, 128
    res = _play(n_games, n_iter, "hash_flat", max_batch=2000)
    assert int((res._soa.n_samples > 0).sum()) == n_games
    exp = oracle.self_play_parallel([(i, 0, 0) for i in range(n_check)], n_iter, C_EXPL, C_PEN, "hash_flat", chunk=8)
    for i in range(n_check):
        assert _records(res._soa, i) == exp[i], f"game_id {i}"


def test_tree_state_64_games_600_sims_with_cache_and_speculation():
    """Tree dumps against the oracle's pointer trees along whole games at n = 600 (every 24th tick for
    all 64 games, every tick for four of them), with the cache, speculative rows and default-size arenas."""
    _need_gpu()
    from c4a0_b200 import _lib as L
    from c4a0_b200.engine import Engine

    n_games, n_iter = 64, 600
    e = Engine(2 * n_games, n_games, n_iter, C_EXPL, C_PEN, flags=L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE,
               arena_blocks=8 * (n_iter + 2), spec_rows=n_games)
    R = e.io_rows
    io = [torch.zeros(R, 2, 6, 7, device="cuda"), torch.zeros(R, 7, device="cuda"), torch.zeros(R, device="cuda"),
          torch.zeros(R, device="cuda")]
    e.bind_io(*[t.data_ptr() for t in io])
    ids = [3 * i for i in range(n_games)]
    e.set_requests(ids, [0] * n_games, [0] * n_games)
    games = [oracle.Game(game_id=g) for g in ids]
    done = [False] * n_games
    checked = 0
    for tick in range(200000):
        e.eval_builtin(L.EVAL_HASH_FLAT)
        e.step()
        for s in range(n_games):
            if done[s] or not (s < 4 or tick % 24 == s % 24):
                continue
            info = e.slot_info(s)
            if info.state == 0:
                done[s] = True
                continue
            g = games[s]
            guard = 0
            while (g.n_moves(), g.root_visit_count()) != (info.n_moves, info.root_visits):
                if g.root_visit_count() >= n_iter:
                    ply = bin(g.root_pos().mask).count("1")
                    assert g.make_random_move(C_EXPL, 4.0 if ply < 4 else (2.0 if ply < 8 else 1.0))
                else:
                    pol, qp, qn = oracle.builtin_eval("hash_flat", g.leaf_pos())
                    g.on_received_policy(pol, qp, qn, C_EXPL, C_PEN)
                guard += 1
                assert guard < 60 * n_iter, "oracle and engine diverged in (moves, visits)"
            assert g.root_pos().key() == (info.root_mask, info.root_value)
            got, exp = e.dump_tree(s), g.dump_tree()
            assert got.size == exp.size and np.array_equal(got, exp), f"tick {tick} slot {s}"
            checked += 1
        if tick % 64 == 63 and e.poll().n_finished == n_games:
            break
    st = e.stats()
    assert e.poll().n_finished == n_games and checked > 1000
    assert st["cache_hits"] > 0 and st["spec_rows"] > 0
    e.close()
