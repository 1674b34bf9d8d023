"""GPU tests of the rows SURVEY.md §8(f) marks "next": device-side sample export with flip
augmentation (N1), the generation trainer (N2), plus edge cases of the engine."""

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _hash_eval_session(n_games, n_slots, n_iter):
    from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession

    g = torch.Generator().manual_seed(1)
    W = (torch.randn(84, generator=g) * 2).cuda()
    V = torch.randn(84, generator=g).cuda()

    def net(planes):
        x = planes[:, :84].float()
        return (x * W).view(-1, 7, 12).sum(2), torch.tanh((x * V).sum(1)), torch.tanh((x * V).sum(1)) * 0.5

    ev = DeviceEvaluator(net, torch.float32, 96)
    sess = SelfPlaySession(n_slots, n_games, n_iter, 6.6, 0.01, plane_dtype=torch.float32, plane_stride=96, n_lanes=1)
    ids = np.arange(n_games) + 7
    z = np.zeros(n_games, np.uint64)
    out, info = sess.play(ids, z, z, ev)
    return sess, ids, out


def test_device_export_matches_host_arrays_and_flip():
    import c4a0_rust as R

    sess, ids, out = _hash_eval_session(300, 128, 20)
    meta = np.stack([ids, ids * 0, ids * 0], axis=1).astype(np.uint64)
    res = R.PlayGamesResult._from_soa(meta, out)
    pos, pol, qp, qn = res.to_arrays()
    dpos, dpol, dqp, dqn = [t.cpu().numpy() for t in sess.export_tensors(augment=True)]
    S = len(pos)
    assert dpos.shape == (2 * S, 2, 6, 7) and dpol.shape == (2 * S, 7)
    assert np.array_equal(dpos[:S], pos) and np.array_equal(dpol[:S], pol)
    assert np.array_equal(dqp[:S], qp) and np.array_equal(dqn[:S], qn)
    # mirror images: Sample.flip_h (types.rs:115-122)
    assert np.array_equal(dpos[S:], pos[:, :, :, ::-1]) and np.array_equal(dpol[S:], pol[:, ::-1])
    assert np.array_equal(dqp[S:], qp) and np.array_equal(dqn[S:], qn)
    s0 = res.results[0].samples[0].flip_h().to_numpy()
    assert np.array_equal(dpos[S], s0[0])
    plain = sess.export_tensors(augment=False)
    assert plain[0].shape[0] == S
    sess.close()


@pytest.mark.parametrize("n_games,n_slots,n_iter,c", [(1, 1, 1, 1.0), (5, 1, 3, 0.0), (3, 3, 2, 6.6), (40, 7, 1, 2.0), (2, 2, 300, 4.0)])
def test_edge_cases_match_oracle(n_games, n_slots, n_iter, c):
    """Single game, single slot (max_nn_batch_size = 1), n_mcts_iterations = 1, c_exploration = 0,
    64-bit game ids — whole records vs the oracle (hash evaluator)."""
    from c4a0_b200 import _lib as L
    from c4a0_b200.engine import Engine

    ids = [2**63 + 11 * i for i in range(n_games)]
    e = Engine(n_slots, n_games, n_iter, c, 0.02)
    io = [torch.zeros(n_slots, 2, 6, 7, device="cuda"), torch.zeros(n_slots, 7, device="cuda"), torch.zeros(n_slots, device="cuda"), torch.zeros(n_slots, device="cuda")]
    e.bind_io(*[t.data_ptr() for t in io])
    e.set_requests(ids, [3] * n_games, [3] * n_games)
    try:
        for i in range(400000):
            e.eval_builtin(L.EVAL_HASH)
            e.step()
            if i % 8 == 7 and e.poll().n_finished == n_games:
                break
    except L.EngineError as exc:
        # n = 1: the reference can sample an unvisited / illegal child and panics (mcts.rs:190-197)
        assert n_iter == 1 and exc.code == L.E_ENGINE
        e.close()
        return
    got = e.fetch_results()
    exp = oracle.self_play([(g, 3, 3) for g in ids], n_slots, n_iter, c, 0.02, evaluator="hash").records()
    for g in range(n_games):
        rec = [
            (int(got.mask[g, k]), int(got.value[g, k]), tuple(got.policy[g, k].view(np.uint32).tolist()),
             int(got.q_penalty[g, k].view(np.uint32)), int(got.q_no_penalty[g, k].view(np.uint32)))
            for k in range(int(got.n_samples[g]))
        ]
        assert rec == exp[g], f"game {g}"
    e.close()


def test_bf16_fused_network_runs_are_reproducible():
    """Row order of the network batch depends on arrival order; results must not."""
    import c4a0_rust as R
    from c4a0_b200.nn import ConnectFourNet, default_config
    from c4a0_b200.selfplay import DeviceEvaluator

    torch.manual_seed(1337)
    model = ConnectFourNet(default_config()).cuda().eval()
    ev = DeviceEvaluator.from_model(model, torch.bfloat16)
    reqs = [R.GameMetadata(i, 0, 0) for i in range(700)]
    # Speculative rows are handed out first come, first served when their budget runs out, so the SIZE and the
    # composition of a batch differ between runs.  The library's network kernel accumulates every output element
    # in a fixed order whatever the batch (tests/test_gpu_net.py), so the shipped default — bf16, evaluation cache
    # and speculation on — must give identical games run after run (round 1's cuBLASLt chain did not).
    assert type(ev).__name__ == "NativeEvaluator"
    a = R.play_games(reqs, 512 + 512, 40, 6.6, 0.01, ev)
    b = R.play_games(reqs, 512 + 512, 40, 6.6, 0.01, ev)
    assert a._run_info.stats["spec_rows"] > 0
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(a._soa, f), getattr(b._soa, f)), f
    st = a._run_info.stats  # equal leaves share rows (rows nobody asked for are speculative)
    assert st["nn_evals"] - st["spec_rows"] < st["leaf_requests"]


def test_training_loop_one_generation(tmp_path):
    """training_test.py:42-82: 1 generation on tiny settings; the saved model differs from its parent,
    val_loss equals the loss computed by hand, artefacts load back."""
    from c4a0_b200 import training as T
    from c4a0_b200.nn import ModelConfig

    cfg = ModelConfig(n_residual_blocks=1, conv_filter_size=4, n_policy_layers=2, n_value_layers=2, lr_schedule={0: 2e-3}, l2_reg=4e-4)
    base = str(tmp_path / "training")
    gen = T.training_loop(base, torch.device("cuda", 0), n_self_play_games=8, n_mcts_iterations=4, c_exploration=6.6,
                          c_ply_penalty=0.01, self_play_batch_size=8, training_batch_size=16, model_config=cfg, max_gens=1,
                          max_epochs=4, nn_dtype=torch.float32, log=None)
    assert gen.gen_n == 1 and gen.val_loss is not None
    gens = T.TrainingGen.load_all(base)
    assert [g.gen_n for g in gens] == [1, 0] and gens[0].parent == gens[1].created_at
    child, parent = gens[0].get_model(base), gens[1].get_model(base)
    assert any(not torch.equal(a, b) for a, b in zip(child.state_dict().values(), parent.state_dict().values()))
    games = gens[0].get_games(base)
    assert len(games.results) == 8 and all(len(g.samples) >= 7 for g in games.results)
    _, val = T.split_arrays(games)
    with torch.no_grad():
        manual = float(T.loss_terms(child.eval(), *[torch.from_numpy(a) for a in val])[0])
    assert abs(manual - gen.val_loss) < 1e-4


def test_multi_model_fast_path_equals_callback_path():
    """Tournament request list (three models): the device fast path with a MultiModelEvaluator plays
    the same games as the numpy-callback path serving the same three networks (N4)."""
    import c4a0_rust as R
    from c4a0_b200.selfplay import DeviceEvaluator, MultiModelEvaluator

    nets = {}
    for mid in (0, 5, 2**63 + 9):
        g = torch.Generator().manual_seed(mid % 1000 + 1)
        W = (torch.randn(84, generator=g) * 2).cuda()
        V = torch.randn(84, generator=g).cuda()

        def net(planes, W=W, V=V):
            x = planes[:, :84].float()
            q = torch.tanh((x * V).sum(1))
            return (x * W).view(-1, 7, 12).sum(2), q, q * 0.25

        nets[mid] = net
    ids = list(nets)
    reqs = [R.GameMetadata(100 + i, ids[i % 3], ids[(i + 1) % 3]) for i in range(90)]

    def cb(model_id, pos):
        with torch.no_grad():
            pol, a, b = nets[model_id](torch.from_numpy(pos).cuda().reshape(len(pos), 84))
        return pol.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy()

    slow = R.play_games(reqs, 64, 16, 3.0, 0.01, cb)
    fast = R.play_games(reqs, 64, 16, 3.0, 0.01, MultiModelEvaluator({m: DeviceEvaluator(f, torch.float32, 96) for m, f in nets.items()}))
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(slow._soa, f), getattr(fast._soa, f)), f
    assert {g.player0_score() for g in fast.results} <= {0.0, 0.5, 1.0}
    with pytest.raises(ValueError):
        R.play_games(reqs, 64, 16, 3.0, 0.01, MultiModelEvaluator({0: DeviceEvaluator(nets[0], torch.float32, 96)}))


def test_gather_session_samples_equals_fetch_results():
    """dist.gather_session_samples packs the valid samples on the device straight from the engine's store;
    unpacked (on the device or on the host) they equal what fetch_results copies."""
    from c4a0_b200 import dist as D
    from c4a0_b200.selfplay import BuiltinEvaluator, SelfPlaySession

    n = 300
    sess = SelfPlaySession(256, n, 40, 6.6, 0.01, eval_cache=True)
    ids = np.arange(n, dtype=np.uint64) * 5 + 2
    z = np.zeros(n, np.uint64)
    out, info = sess.play(ids, z, z, BuiltinEvaluator("hash"))
    meta = np.stack([ids, z, z], axis=1)
    gm, gs = D.gather_session_samples(sess, meta)
    assert np.array_equal(gm, meta)
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(gs, f), getattr(out, f)), f
    gm2, counts, packed = D.gather_session_samples(sess, meta, packed_result=True)
    assert packed.shape == (int(out.n_samples.sum()), D.PACK_WORDS) and np.array_equal(counts, out.n_samples.astype(np.int32))
    again = D.unpack_samples(counts, packed)
    for f in ("mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(again, f), getattr(out, f)), f
    sess.close()
