"""GPU tests of the drop-in boundary: `c4a0_rust.play_games` and its result classes.

Modelled on the reference's own Python tests (tests/c4a0_tests/pybridge_test.py:22-39,
tournament_test.py:27-51) and on rust/src/self_play.rs:405-459.
"""

import pickle

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def uniform_cb(model_id, pos):
    assert pos.dtype == np.float32 and pos.shape[1:] == (2, 6, 7) and pos.flags["C_CONTIGUOUS"]
    n = len(pos)
    return np.ones((n, 7), np.float32) / 7, np.zeros(n, np.float32), np.zeros(n, np.float32)


def test_play_games_numpy_callback_like_pybridge_test():
    import c4a0_rust as R

    reqs = [R.GameMetadata(i, 0, 0) for i in range(4)]
    res = R.play_games(reqs, 8, 2, 1.4, 0.01, uniform_cb)
    assert len(res.results) == 4
    for g in res.results:
        assert len(g.samples) >= 7
        pos, pol, qp, qn = g.samples[0].to_numpy()
        assert pos.shape == (2, 6, 7) and pol.shape == (7,) and qp.shape == () and qn.shape == ()
        assert g.player0_score() in (0.0, 0.5, 1.0)
    # split_train_test is deterministic and does not mutate (pybridge_test.py:33-39)
    a = res.split_train_test(0.5, 1337)
    b = res.split_train_test(0.5, 1337)
    assert a == b and len(a[0]) + len(a[1]) == sum(len(g.samples) for g in res.results)
    # pickle round trip through CBOR
    again = pickle.loads(pickle.dumps(res))
    assert [g.samples for g in again.results] == [g.samples for g in res.results]
    assert (res + again).unique_positions() == res.unique_positions()


def test_self_play_invariants_like_reference_test():
    """self_play.rs:405-459: one game, 50 iterations, uniform evaluator with batch <= 10."""
    import c4a0_rust as R

    def cb(model_id, pos):
        assert len(pos) <= 10
        return uniform_cb(model_id, pos)

    res = R.play_games([R.GameMetadata(0, 0, 0)], 10, 50, 1.0, 0.01, cb)
    (g,) = res.results
    assert len(g.samples) >= 7
    planes = [s.to_numpy()[0] for s in g.samples]
    assert sum(1 for p in planes if p.sum() == 0) == 1  # exactly one start position
    terminal = [s for s in g.samples if oracle.terminal_state(oracle.Pos(s._mask, s._value)) != 0]
    assert len(terminal) == 1
    assert float(terminal[0].to_numpy()[3]) in (-1.0, 0.0, 1.0)


def test_callback_path_matches_oracle_with_hash_network_and_two_models():
    """Tournament-style request list (player0 != player1): per game_id records equal the oracle's
    when both use the same deterministic evaluator keyed by (model, position)."""
    import c4a0_rust as R

    def net(model_id, keys):
        pol = np.zeros((len(keys), 7), np.float32)
        qp = np.zeros(len(keys), np.float32)
        qn = np.zeros(len(keys), np.float32)
        for i, (m, v) in enumerate(keys):
            p, a, b = oracle.builtin_eval("hash", oracle.Pos(m, v), model_id)
            pol[i], qp[i], qn[i] = p, a, b
        return pol, qp, qn

    def cb(model_id, pos):
        # rebuild (mask, value) from the planes the engine produced — this also checks the planes
        n = len(pos)
        flat = pos.reshape(n, 2, 42)
        w = (np.uint64(1) << np.arange(42, dtype=np.uint64))[None, :]
        mine = (flat[:, 0].astype(np.uint64) * w).sum(1)
        theirs = (flat[:, 1].astype(np.uint64) * w).sum(1)
        keys = [(int(a | b), int(a)) for a, b in zip(mine, theirs)]
        assert len(set(keys)) == len(keys), "positions in one batch must be unique (self_play.rs:203-208)"
        return net(model_id, keys)

    reqs = [(100 + i, i % 3, (i + 1) % 3) for i in range(12)]
    res = R.play_games([R.GameMetadata(*r) for r in reqs], 5, 24, 3.0, 0.01, cb)
    exp = oracle.self_play(reqs, 5, 24, 3.0, 0.01, evaluator=net)
    for i, g in enumerate(res.results):
        got = [(s._mask, s._value, s._policy.tobytes(), s._q_penalty.tobytes(), s._q_no_penalty.tobytes()) for s in g.samples]
        want = [
            (int(s.pos.mask), int(s.pos.value), np.array(list(s.policy), np.float32).tobytes(),
             np.float32(s.q_penalty).tobytes(), np.float32(s.q_no_penalty).tobytes())
            for s in exp.samples[i]
        ]
        assert got == want, f"game {i}"
        assert g.player0_score() == oracle.player0_score(exp.samples[i])


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_device_fast_path_real_network_matches_oracle_given_same_outputs(dtype):
    """Tier E2: the real ResNet on the GPU fast path (CUDA graph, zero-copy planes).  The engine's
    leaf positions and the network outputs it consumed are memoised per tick; the oracle then plays
    the same games reading the memo, so both see identical NN outputs by construction."""
    import c4a0_rust as R
    from c4a0_b200.nn import ConnectFourNet, ModelConfig
    from c4a0_b200.selfplay import DeviceEvaluator

    dt = getattr(torch, dtype)
    torch.manual_seed(1337)
    model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=8, n_policy_layers=2, n_value_layers=2))
    model = model.to("cuda", dt).eval()
    memo = {}

    n_games, n_iter = 24, 32
    from c4a0_b200.selfplay import SelfPlaySession

    sess = SelfPlaySession(n_games, n_games, n_iter, 6.6, 0.01, plane_dtype=dt, n_lanes=1)
    ids = np.arange(500, 500 + n_games)
    zeros = np.zeros(n_games, np.uint64)

    def evaluator(planes):
        return model(planes.view(-1, 2, 6, 7))

    # drive tick by tick so that the (leaf -> outputs) pairs can be recorded
    with torch.cuda.stream(sess.stream):
        s = sess.stream.cuda_stream
        sess.engine.set_requests(ids, zeros, zeros, s)
        for tick in range(100000):
            n_rows, mask, value, _ = sess.engine.fetch_rows(s)
            with torch.no_grad():
                pol, a, b = evaluator(sess.planes)
                sess.logits.copy_(pol)
                sess.qp.copy_(a)
                sess.qn.copy_(b)
            sess.stream.synchronize()
            lg, qa, qb = sess.logits.cpu().numpy(), sess.qp.cpu().numpy(), sess.qn.cpu().numpy()
            assert len(set(zip(mask.tolist(), value.tolist()))) == n_rows  # rows are distinct positions
            for r in range(n_rows):
                key = (int(mask[r]), int(value[r]))
                # the memo is authoritative: the first answer for a position is what every later
                # visit (in either engine) consumes, whatever row / batch it was computed in
                rec = memo.setdefault(key, (lg[r].tobytes(), qa[r].tobytes(), qb[r].tobytes()))
                lg[r] = np.frombuffer(rec[0], np.float32)
                qa[r] = np.frombuffer(rec[1], np.float32)[0]
                qb[r] = np.frombuffer(rec[2], np.float32)[0]
            sess.logits.copy_(torch.from_numpy(lg))
            sess.qp.copy_(torch.from_numpy(qa))
            sess.qn.copy_(torch.from_numpy(qb))
            sess.engine.step(s)
            if sess.engine.poll(s).n_finished == n_games:
                break
        got = sess.engine.fetch_results(0, n_games, s)
    sess.close()

    def memo_net(model_id, keys):
        # The reference (and so the oracle) also sends TERMINAL leaves to the network and ignores
        # the answer (mcts.rs:92-98); the engine never does, so those are not in the
        # memo.  Any other missing key means the two searches visited different positions.
        pol = np.zeros((len(keys), 7), np.float32)
        qp = np.zeros(len(keys), np.float32)
        qn = np.zeros(len(keys), np.float32)
        for i, k in enumerate(keys):
            if k in memo:
                pol[i] = np.frombuffer(memo[k][0], np.float32)
                qp[i] = np.frombuffer(memo[k][1], np.float32)[0]
                qn[i] = np.frombuffer(memo[k][2], np.float32)[0]
            else:
                assert oracle.terminal_state(oracle.Pos(*k)) != 0, f"oracle visited {k}, the engine did not"
        return pol, qp, qn

    exp = oracle.self_play([(int(i), 0, 0) for i in ids], n_games, n_iter, 6.6, 0.01, evaluator=memo_net)
    mismatched = 0
    for i in range(n_games):
        n = int(got.n_samples[i])
        assert n == len(exp.samples[i])
        for k in range(n):
            s = exp.samples[i][k]
            assert (int(got.mask[i, k]), int(got.value[i, k])) == (int(s.pos.mask), int(s.pos.value))
            np.testing.assert_allclose(got.policy[i, k], np.array(list(s.policy), np.float32), atol=1e-3)
            assert abs(got.q_penalty[i, k] - s.q_penalty) <= 1e-3
            mismatched += got.policy[i, k].tobytes() != np.array(list(s.policy), np.float32).tobytes()
    assert mismatched == 0  # in practice 0 ulp, not just 1e-3


def test_fast_path_through_play_games_native_loop_equals_python_loop():
    import c4a0_rust as R
    from c4a0_b200 import selfplay
    from c4a0_b200.nn import ConnectFourNet, ModelConfig

    torch.manual_seed(7)
    model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=8, n_policy_layers=2, n_value_layers=2))
    model = model.to("cuda").eval()
    reqs = [R.GameMetadata(i, 0, 0) for i in range(40)]
    old = dict(selfplay.DEFAULTS)
    try:
        selfplay.DEFAULTS.update(host_loop="native")
        a = R.play_games(reqs, 16, 20, 6.6, 0.01, model)  # 40 games through 16 slots: refill
        selfplay.DEFAULTS.update(host_loop="python", poll_every=8)
        b = R.play_games(reqs, 16, 20, 6.6, 0.01, model)
    finally:
        selfplay.DEFAULTS.update(old)
    assert a._run_info.stats["samples"] == b._run_info.stats["samples"] > 0
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(a._soa, f), getattr(b._soa, f)), f
    assert all(len(g.samples) >= 7 for g in a.results)


def test_errors_are_exceptions_not_aborts():
    import c4a0_rust as R

    with pytest.raises(TypeError):
        R.play_games([1, 2, 3], 8, 2, 1.0, 0.01, uniform_cb)
    with pytest.raises(ValueError):
        R.play_games([R.GameMetadata(0, 0, 0)], 8, 2, 1.0, 0.01, lambda m, p: (np.zeros((1, 3), np.float32), 0, 0))
    with pytest.raises(ValueError):
        R.PlayGamesResult.from_cbor(b"\x01\x02")
    with pytest.raises(NotImplementedError):
        R.run_tui(uniform_cb, 10, 1.0, 0.01)
    assert len(R.play_games([], 8, 2, 1.0, 0.01, uniform_cb).results) == 0
