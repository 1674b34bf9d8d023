"""GPU tests of the evaluation cache (C4A0_FLAG_EVAL_CACHE): answers of the
network are reused for the same (position, model) within one set_requests() job.  The parity tests in
test_gpu_engine.py run every oracle comparison with the cache on as well; here: what it saves, that
it never crosses jobs or models, and the session-level switch."""

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from test_gpu_engine import _make_engine, _need_gpu, _records, _run_builtin  # noqa: E402


def test_cache_saves_rows_and_ticks_not_results():
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 256, 80, 6.6, 0.01
    ids = [11 * i + 3 for i in range(n_games)]
    outs, stats, ticks = [], [], []
    for flags in (0, L.FLAG_EVAL_CACHE):
        e, io = _make_engine(n_games, n_games, n_iter, c_expl, c_pen, flags=flags)
        e.set_requests(ids, [0] * n_games, [0] * n_games)
        ticks.append(_run_builtin(e, 0, poll_every=1))  # uniform priors: broad trees, many transpositions
        outs.append(e.fetch_results())
        stats.append(e.stats())
        e.close()
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(outs[0], f), getattr(outs[1], f)), f
    a, b = stats
    for k in ("sims", "terminal_leaf_sims", "moves", "samples", "expansions", "select_depth_sum", "skipped_root_sims"):
        assert a[k] == b[k], k
    assert a["cache_hits"] == 0 and a["cache_inserts"] == 0
    # every expansion is answered by the network or by the cache
    assert b["cache_hits"] + b["leaf_requests"] == b["expansions"] == a["leaf_requests"]
    assert b["cache_hits"] > 0.05 * b["expansions"]
    assert b["nn_evals"] < a["nn_evals"]
    assert ticks[1] < ticks[0]
    assert b["cache_inserts"] <= b["nn_evals"]


def test_cache_does_not_cross_jobs_or_models():
    """Job 1 fills the cache from the hash evaluator; job 2 on the same engine plays the same ids
    with the uniform evaluator and must match the uniform oracle.  Then a two-model job: the hash
    evaluator depends on the model id, so a hit across models would change the games."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 48, 40, 6.6, 0.01
    reqs = [(100 + i, 0, 0) for i in range(n_games)]
    e, io = _make_engine(n_games, n_games, n_iter, c_expl, c_pen, flags=L.FLAG_EVAL_CACHE, eval_cache_entries=1 << 16)
    for kind, name in ((1, "hash"), (0, "uniform"), (1, "hash")):
        e.set_requests([r[0] for r in reqs], [0] * n_games, [0] * n_games)
        _run_builtin(e, kind)
        got = e.fetch_results()
        exp = oracle.self_play(reqs, n_games, n_iter, c_expl, c_pen, evaluator=name).records()
        for i in range(n_games):
            assert _records(got, i) == exp[i], (name, i)
        assert e.stats()["cache_hits"] > 0
    reqs2 = [(100 + i, 7, 9) if i % 2 else (100 + i, 9, 7) for i in range(n_games)]
    e.set_requests([r[0] for r in reqs2], [r[1] for r in reqs2], [r[2] for r in reqs2])
    _run_builtin(e, 1)
    got = e.fetch_results()
    exp = oracle.self_play(reqs2, n_games, n_iter, c_expl, c_pen, evaluator="hash").records()
    for i in range(n_games):
        assert _records(got, i) == exp[i], i
    e.close()


def test_play_games_fast_path_uses_the_cache_and_callbacks_do_not():
    _need_gpu()
    import c4a0_rust
    from c4a0_b200 import selfplay
    from c4a0_b200.selfplay import DeviceEvaluator

    g = torch.Generator().manual_seed(5)
    W = (torch.randn(84, generator=g) * 2).cuda()
    V = torch.randn(84, generator=g).cuda()

    def net(planes):  # per-row output independent of the batch it is computed in
        x = planes[:, :84].float()
        pol = (x * W).view(-1, 7, 12).sum(2)
        q = torch.tanh((x * V).sum(1))
        return pol, q, q * 0.5

    def cb(model_id, pos):
        with torch.no_grad():
            pol, a, b = net(torch.from_numpy(pos).cuda().reshape(len(pos), 84))
        return pol.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy()

    reqs = [c4a0_rust.GameMetadata(3 * i + 1, 0, 0) for i in range(300)]
    ev = DeviceEvaluator(net, torch.float32, 96)
    res = {}
    try:
        for mode in ("cache", "nocache"):
            selfplay.DEFAULTS["eval_cache"] = mode == "cache"
            res[mode] = c4a0_rust.play_games(reqs, 300, 30, 6.6, 0.01, ev)
        res["callback"] = c4a0_rust.play_games(reqs, 300, 30, 6.6, 0.01, cb)
    finally:
        selfplay.DEFAULTS["eval_cache"] = True
        c4a0_rust._native.close_cached_session()
    assert res["cache"]._run_info.stats["cache_hits"] > 0
    assert res["nocache"]._run_info.stats["cache_hits"] == 0
    assert res["callback"]._run_info.stats["cache_hits"] == 0
    st = res["cache"]._run_info.stats  # rows games asked for = all rows - the speculative ones
    assert st["nn_evals"] - st["spec_rows"] < res["nocache"]._run_info.stats["nn_evals"]
    a = res["cache"].to_arrays()
    for other in ("nocache", "callback"):
        b = res[other].to_arrays()
        for x, y in zip(a, b):
            assert np.array_equal(x, y), other


def test_speculation_fills_spare_rows_and_saves_ticks():
    """256 games in 1024 slots: every batch has spare rows; the children of expanded leaves are
    evaluated in them, so later leaves are already answered.  Records stay those of the plain engine."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_slots, n_iter, c_expl, c_pen = 256, 1024, 80, 6.6, 0.01
    ids = [11 * i + 3 for i in range(n_games)]
    outs, stats, ticks = [], [], []
    for flags in (0, L.FLAG_EVAL_CACHE, L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE):
        e, io = _make_engine(n_slots, n_games, n_iter, c_expl, c_pen, flags=flags, spec_rows=768)
        e.set_requests(ids, [0] * n_games, [0] * n_games)
        seen_max = 0
        for t in range(200000):
            e.eval_builtin(1)
            e.step()
            p = e.poll()
            seen_max = max(seen_max, p.n_rows)
            if p.n_finished == p.n_requests:
                break
        ticks.append(t + 1)
        assert seen_max <= e.io_rows
        outs.append(e.fetch_results())
        stats.append(e.stats())
        e.close()
    for other in (1, 2):
        for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
            assert np.array_equal(getattr(outs[0], f), getattr(outs[other], f)), (other, f)
    plain, cache, spec = stats
    for k in ("sims", "terminal_leaf_sims", "moves", "samples", "expansions", "select_depth_sum"):
        assert plain[k] == cache[k] == spec[k], k
    assert plain["spec_rows"] == 0 and cache["spec_rows"] == 0 and spec["spec_rows"] > 0
    assert spec["cache_hits"] > cache["cache_hits"]
    assert spec["cache_hits"] + spec["leaf_requests"] == spec["expansions"]
    assert ticks[2] < ticks[1] < ticks[0]


def test_speculate_needs_the_cache():
    _need_gpu()
    from c4a0_b200 import _lib as L

    with pytest.raises(L.EngineError):
        _make_engine(8, 8, 10, 1.0, 0.01, flags=L.FLAG_SPECULATE)
