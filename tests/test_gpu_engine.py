"""GPU parity tests: the CUDA engine (through the C-ABI) against the CPU oracle.

Integer work (bitboards, legality, terminal detection, tree structure, sampled moves, game
records) must be bit-exact; the float fields are compared by bit pattern too (tolerance 0 — the
north star allows 1e-3, the design gives 0 ulp, so any drift is a bug worth seeing).
"""

import numpy as np
import pytest

import oracle
from oracle import Pos

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def random_positions(n, seed):
    """The reference's proptest strategy random_pos() (c4r.rs:610-629)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        k = int(rng.integers(0, 60))
        out.append(oracle.random_pos(rng.integers(0, 7, size=k).tolist()))
    return out


# ------------------------------------------------------------------------------------------------
def test_rules_kernel_bit_exact():
    _need_gpu()
    from c4a0_b200 import engine as E

    ps = random_positions(20000, 1)
    ps.append(Pos(0, 0))
    m = np.array([p.mask for p in ps], np.uint64)
    v = np.array([p.value for p in ps], np.uint64)
    out = E.rules_batch(m, v, 0.01)
    for i, p in enumerate(ps):
        assert out["terminal"][i] == oracle.terminal_state(p)
        lg = oracle.legal_moves(p)
        assert out["legal"][i] == sum(1 << c for c in range(7) if lg[c])
        assert out["ply"][i] == bin(p.mask).count("1")
        tv = oracle.terminal_value(p, 0.01)
        if tv is not None:
            assert np.float32(tv[0]).tobytes() == out["q_penalty"][i].tobytes()
            assert np.float32(tv[1]).tobytes() == out["q_no_penalty"][i].tobytes()
        for c in range(7):
            ch = oracle.make_move(p, c)
            if ch is None:
                assert out["child_mask"][i, c] == 0 and out["child_value"][i, c] == 0
            else:
                assert (int(out["child_mask"][i, c]), int(out["child_value"][i, c])) == ch.key()
        if i % 16 == 0:
            assert np.array_equal(out["planes"][i], oracle.planes(p))
            f = oracle.lib().c4o_flip_h(p)
            assert (int(out["flip_mask"][i]), int(out["flip_value"][i])) == f.key()


def test_device_logf_expf_match_libm():
    _need_gpu()
    from c4a0_b200 import _lib as L
    from c4a0_b200 import engine as E

    rng = np.random.default_rng(2)
    n = 4_000_000
    bits = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    x = np.concatenate([x, np.arange(0, 70000, dtype=np.float32), -rng.random(1_000_000, dtype=np.float32) * 120])
    for op, ref in ((L.MATH_LOGF, oracle.logf), (L.MATH_EXPF, oracle.expf)):
        got = E.math_batch(op, x)
        exp = ref(x)
        same = (got.view(np.uint32) == exp.view(np.uint32)) | (np.isnan(got) & np.isnan(exp))
        assert same.all(), f"op {op}: {np.count_nonzero(~same)} mismatches, first at x={x[~same][0]!r}"


def test_softmax_and_sampling_kernels():
    _need_gpu()
    from c4a0_b200 import engine as E

    rng = np.random.default_rng(3)
    n = 20000
    logits = (rng.standard_normal((n, 7)) * 3).astype(np.float32)
    legal = rng.integers(1, 128, size=n).astype(np.uint32)
    got = E.softmax_batch(logits, legal)
    for i in range(0, n, 7):
        x = [float(logits[i, c]) if (legal[i] >> c) & 1 else float("-inf") for c in range(7)]
        exp = oracle.softmax(x)
        assert got[i].tobytes() == exp.tobytes()
    # visit-count shaped policies, the three temperatures of the schedule, many seeds
    counts = rng.integers(0, 400, size=(n, 7)).astype(np.float32)
    counts[rng.random((n, 7)) < 0.2] = 0
    counts[counts.sum(1) == 0, 3] = 1
    pol = (counts / counts.sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)
    temp = rng.choice(np.array([4.0, 2.0, 1.0, 0.0], np.float32), size=n)
    seed = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    tempered, col = E.sample_batch(pol, temp, seed)
    for i in range(0, n, 5):
        exp_t = oracle.apply_temperature(pol[i].tolist(), float(temp[i]))
        assert tempered[i].tobytes() == exp_t.tobytes(), (i, pol[i], temp[i])
        assert col[i] == oracle.weighted_index_sample(exp_t.tolist(), int(seed[i]))


# ------------------------------------------------------------------------------------------------
def _make_engine(n_slots, n_req, n_iter, c_expl, c_pen, **kw):
    from c4a0_b200.engine import Engine

    e = Engine(n_slots, n_req, n_iter, c_expl, c_pen, **kw)
    R = e.io_rows  # n_slots (+ the speculative rows)
    io = dict(
        planes=torch.zeros(R, 2, 6, 7, device="cuda"),
        logits=torch.zeros(R, 7, device="cuda"),
        qp=torch.zeros(R, device="cuda"),
        qn=torch.zeros(R, device="cuda"),
    )
    e.bind_io(io["planes"].data_ptr(), io["logits"].data_ptr(), io["qp"].data_ptr(), io["qn"].data_ptr())
    return e, io


def _run_builtin(e, kind, max_steps=200000, poll_every=16):
    for i in range(max_steps):
        e.eval_builtin(kind)
        e.step()
        if i % poll_every == poll_every - 1:
            p = e.poll()
            if p.n_finished == p.n_requests:
                return i + 1
    raise AssertionError("self-play did not finish")


def _records(samples, i):
    n = int(samples.n_samples[i])
    return [
        (
            int(samples.mask[i, k]),
            int(samples.value[i, k]),
            tuple(samples.policy[i, k].view(np.uint32).tolist()),
            int(samples.q_penalty[i, k].view(np.uint32)),
            int(samples.q_no_penalty[i, k].view(np.uint32)),
        )
        for k in range(n)
    ]


@pytest.mark.parametrize("cache", [0, 1 << 20, 32, -1], ids=["nocache", "cache", "cache32", "speculate"])
@pytest.mark.parametrize("kind,name", [(0, "uniform"), (1, "hash")])
def test_tree_state_matches_oracle_every_step(kind, name, cache):
    """Tree structure + every N / Qp / Qn / prior bit pattern after every lockstep tick; also with the
    evaluation cache (several simulations per tick; 32 entries = constant replacement) and with
    speculative evaluation of children in the spare rows (64 slots for 6 games)."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 6, 40, 1.7, 0.01
    spec = cache == -1
    e, io = _make_engine(64 if spec else n_games, n_games, n_iter, c_expl, c_pen, max_inline_sims=3,
                         arena_blocks=70 if kind else 0,
                         flags=(L.FLAG_EVAL_CACHE if cache else 0) | (L.FLAG_SPECULATE if spec else 0),
                         eval_cache_entries=(1 << 16) if spec else cache, spec_rows=48 if spec else 0)
    ids = [0, 1, 2, 77, 1234567, 2**40 + 5]
    e.set_requests(ids, [0] * n_games, [0] * n_games)
    games = [oracle.Game(game_id=g) for g in ids]

    def oracle_sim(g):
        pol, qp, qn = oracle.builtin_eval(name, g.leaf_pos())
        g.on_received_policy(pol, qp, qn, c_expl, c_pen)

    done = [False] * n_games
    for tick in range(4000):
        e.eval_builtin(kind)
        e.step()
        for s in range(n_games):
            if done[s]:
                continue
            info = e.slot_info(s)
            if info.state == 0:  # finished
                done[s] = True
                continue
            g = games[s]
            # advance the oracle game to the same (n_moves, root visits)
            guard = 0
            while (g.n_moves(), g.root_visit_count()) != (info.n_moves, info.root_visits):
                if g.root_visit_count() >= n_iter:
                    ply = bin(g.root_pos().mask).count("1")
                    t = 4.0 if ply < 4 else (2.0 if ply < 8 else 1.0)
                    assert g.make_random_move(c_expl, t)
                else:
                    oracle_sim(g)
                guard += 1
                assert guard < 10 * n_iter, "oracle and engine diverged in (moves, visits)"
            assert g.root_pos().key() == (info.root_mask, info.root_value)
            got = e.dump_tree(s)
            exp = g.dump_tree()
            assert got.size == exp.size and np.array_equal(got, exp), f"tick {tick} slot {s}"
        if all(done):
            break
    assert all(done)
    e.close()


@pytest.mark.parametrize("cache", [0, 0xFFFFFFFF, 256, -1, -256], ids=["nocache", "cache", "cache256", "speculate", "speculate256"])
@pytest.mark.parametrize("kind,name", [(0, "uniform"), (1, "hash")])
@pytest.mark.parametrize("n_slots,arena", [(64, 0), (17, 0), (64, 8 * 62), (33, 100)])
def test_game_records_match_oracle(kind, name, n_slots, arena, cache):
    """Whole games (positions, policies, q values, sampled moves) per game_id, with slot refill,
    with the minimal arena (a compaction at almost every move), a roomy one (none) and one between;
    without and with the evaluation cache (default size, and 256 entries = constant replacement)."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 64, 60, 6.6, 0.01
    reqs = [(1000 + 37 * i, 0, 0) for i in range(n_games)]
    spec = cache < 0
    if spec:  # speculation uses rows no game can ask for: twice the slots the games need
        cache = 0xFFFFFFFF if cache == -1 else -cache
        n_slots = 2 * n_slots
    e, io = _make_engine(n_slots, n_games, n_iter, c_expl, c_pen, arena_blocks=arena,
                         flags=(L.FLAG_EVAL_CACHE if cache else 0) | (L.FLAG_SPECULATE if spec else 0),
                         eval_cache_entries=0 if cache == 0xFFFFFFFF else cache, spec_rows=40 if spec else 0)
    e.set_requests([r[0] for r in reqs], [0] * n_games, [0] * n_games)
    _run_builtin(e, kind)
    got = e.fetch_results()
    exp = oracle.self_play(reqs, n_games, n_iter, c_expl, c_pen, evaluator=name).records()
    for i in range(n_games):
        assert _records(got, i) == exp[i], f"game {i}"
    st = e.stats()
    assert st["samples"] == sum(len(r) for r in exp)
    assert (st["cache_hits"] > 0) == bool(cache) and (st["cache_inserts"] > 0) == bool(cache)
    assert (st["spec_rows"] > 0) == spec
    if arena == 0:
        assert st["compactions"] > n_games and st["compacted_blocks"] > 0
    if arena == 8 * 62 and n_slots == n_games:
        assert st["compactions"] < n_games * 3  # long uniform-evaluator games still fill a half now and then
    e.close()


def test_stats_match_oracle_counts():
    _need_gpu()
    n_games, n_iter, c_expl, c_pen = 32, 50, 6.6, 0.01
    reqs = [(i, 0, 0) for i in range(n_games)]
    e, io = _make_engine(n_games, n_games, n_iter, c_expl, c_pen)
    e.set_requests([r[0] for r in reqs], [0] * n_games, [0] * n_games)
    _run_builtin(e, 1)
    st = e.stats()
    o = oracle.self_play(reqs, n_games, n_iter, c_expl, c_pen, evaluator="hash")
    # the reference counts the sims it wastes on terminal roots; the engine skips them (F9)
    assert st["sims"] + st["skipped_root_sims"] == o.stats["sims"]
    assert st["terminal_leaf_sims"] == o.stats["terminal_leaf_sims"]
    assert st["moves"] == o.stats["moves"]
    assert st["samples"] == o.stats["samples"]
    e.close()


def test_dedup_changes_network_rows_not_results():
    """Equal leaves share a network row (self_play.rs:203-208); game records must not change."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 96, 40, 6.6, 0.01
    ids = [5 * i for i in range(n_games)]
    outs, stats = [], []
    for flags in (0, L.FLAG_NO_DEDUP):
        e, io = _make_engine(n_games, n_games, n_iter, c_expl, c_pen, flags=flags)
        e.set_requests(ids, [0] * n_games, [0] * n_games)
        p = e.poll()
        assert p.n_rows == (1 if flags == 0 else n_games)  # every game starts from the empty board
        _run_builtin(e, 1)
        outs.append(e.fetch_results())
        stats.append(e.stats())
        e.close()
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(outs[0], f), getattr(outs[1], f)), f
    valid = np.arange(43)[None, :] < outs[0].n_samples[:, None]
    assert not outs[0].mask[~valid].any() and not outs[0].policy[~valid].any()  # unused cells are zero
    assert stats[0]["sims"] == stats[1]["sims"] and stats[0]["leaf_requests"] == stats[1]["leaf_requests"]
    assert stats[1]["nn_evals"] == stats[1]["leaf_requests"]
    assert stats[0]["nn_evals"] < stats[0]["leaf_requests"]


def test_rows_are_dense_and_distinct():
    _need_gpu()
    n_games = 200
    e, io = _make_engine(n_games, n_games, 30, 6.6, 0.01)
    e.set_requests(list(range(n_games)), [0] * n_games, [1] * n_games)
    for tick in range(400):
        e.eval_builtin(1)
        e.step()
        if tick % 25 == 0:
            n_rows, mask, value, model = e.fetch_rows()
            keys = list(zip(mask.tolist(), value.tolist(), model.tolist()))
            assert len(set(keys)) == n_rows
            # every waiting game points at a live row, and every live row is used
            used = set()
            for s in range(n_games):
                info = e.slot_info(s)
                if info.state == 1:
                    assert info.nn_row < n_rows
                    used.add(info.nn_row)
            assert used == set(range(n_rows))
    e.close()


def test_native_host_loop_two_lanes_equals_python_loop_one_lane():
    """c4a0_engine_run (bucketed CUDA graphs, two engines on two streams) plays the same games as
    the Python loop on one engine, for an evaluator whose per-row output does not depend on the
    batch it is computed in."""
    _need_gpu()
    from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession

    g = torch.Generator().manual_seed(3)
    W = (torch.randn(84, generator=g) * 2).cuda()
    V = torch.randn(84, generator=g).cuda()

    def net(planes):
        x = planes[:, :84].float()
        pol = (x * W).view(-1, 7, 12).sum(2)
        q = torch.tanh((x * V).sum(1))
        return pol, q, q * 0.5

    ev = DeviceEvaluator(net, torch.float32, 96)
    n_games, n_iter = 1500, 24
    ids = np.arange(n_games) * 3 + 1
    z = np.zeros(n_games, np.uint64)
    res = []
    for lanes, loop, slots in ((2, "native", 1100), (1, "python", 700)):
        sess = SelfPlaySession(slots, n_games, n_iter, 6.6, 0.01, plane_dtype=torch.float32, plane_stride=96, n_lanes=lanes)
        out, info = sess.play(ids, z, z, ev, host_loop=loop, poll_every=16, sample_kernels_every=7)
        assert info.stats["samples"] == int(out.n_samples.sum()) and (out.n_samples >= 8).all()
        assert info.kernel_ms["samples"] > 0 and info.device_s > 0
        res.append(out)
        sess.close()
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(res[0], f), getattr(res[1], f)), f


_BIG_ORACLE = {}


@pytest.mark.parametrize("cache", [0, 1], ids=["nocache", "cache"])
def test_more_games_than_one_wave_of_k_step(cache):
    """40,000 resident games do not fit the GPU at once (one wave of k_step holds ~16,500): warps that
    start late in a tick still have to find the network row of a leaf whose leading game, in an earlier
    wave, has already published its next leaf."""
    _need_gpu()
    from c4a0_b200 import _lib as L

    n_games, n_iter, c_expl, c_pen = 40000, 6, 6.6, 0.01
    reqs = [(7 * i + 1, 0, 0) for i in range(n_games)]
    e, io = _make_engine(n_games, n_games, n_iter, c_expl, c_pen, flags=L.FLAG_EVAL_CACHE if cache else 0)
    e.set_requests([r[0] for r in reqs], [0] * n_games, [0] * n_games)
    _run_builtin(e, 1, poll_every=8)
    got = e.fetch_results()
    key = (n_games, n_iter)
    if key not in _BIG_ORACLE:
        exp = oracle.self_play(reqs, n_games, n_iter, c_expl, c_pen, evaluator="hash")
        _BIG_ORACLE[key] = (exp.stats, exp.records())
    ostats, rec = _BIG_ORACLE[key]
    st = e.stats()
    assert st["samples"] == ostats["samples"] and st["moves"] == ostats["moves"]
    assert st["nn_evals"] < st["leaf_requests"]  # equal leaves shared rows: there were followers
    for i in range(0, n_games, 7):
        assert _records(got, i) == rec[i], f"game {i}"
    e.close()
