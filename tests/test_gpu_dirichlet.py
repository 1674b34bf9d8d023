"""Dirichlet noise on root priors (c4a0_config.dirichlet_alpha / dirichlet_epsilon; BASELINE north star).

The reference has no such noise (rust/src/mcts.rs:114-132 stores the masked softmax unchanged), so there is no
oracle to compare with: the option is OFF by default and these tests pin what it promises —
  * off (alpha = 0 or epsilon = 0) is bit-identical to the parity path;
  * with epsilon = 1 the root's priors ARE the Dirichlet draw: non-negative, summing to one over the legal
    moves, with the mean and variance of Dir(alpha);
  * the draw is a pure function of (game_id, moves played, column): game records are reproducible and do not
    depend on slots, refill order, the evaluation cache or speculation;
  * only roots are touched: a node's priors change when it becomes the root of a search, never deeper.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

C_EXPL, C_PEN = 6.6, 0.01


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _engine(n_slots, n_games, n_iter, flags=0, alpha=0.0, eps=0.0, spec_rows=0):
    from c4a0_b200 import _lib as L
    from c4a0_b200.engine import Engine

    e = Engine(n_slots, n_games, n_iter, C_EXPL, C_PEN, flags=flags, spec_rows=spec_rows, dirichlet_alpha=alpha,
               dirichlet_epsilon=eps)
    R = e.io_rows
    io = (torch.zeros(R, 2, 6, 7, device="cuda"), torch.zeros(R, 7, device="cuda"), torch.zeros(R, device="cuda"),
          torch.zeros(R, device="cuda"))
    e.bind_io(*[t.data_ptr() for t in io])
    return e, io, L


def _play(n_slots, n_games, n_iter, flags=0, alpha=0.0, eps=0.0, spec_rows=0, ids=None):
    e, io, L = _engine(n_slots, n_games, n_iter, flags, alpha, eps, spec_rows)
    ids = list(range(n_games)) if ids is None else ids
    e.set_requests(ids, [0] * n_games, [0] * n_games)
    for i in range(200000):
        e.eval_builtin(L.EVAL_HASH)
        e.step()
        if i % 16 == 15 and e.poll().n_finished == n_games:
            break
    got = e.fetch_results()
    e.close()
    return got


def _same(a, b):
    return all(np.array_equal(getattr(a, f).view(np.uint8), getattr(b, f).view(np.uint8))
               for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"))


def test_off_is_the_parity_path():
    _need_gpu()
    base = _play(32, 48, 48)
    assert _same(base, _play(32, 48, 48, alpha=0.0, eps=0.25))
    assert _same(base, _play(32, 48, 48, alpha=0.3, eps=0.0))


@pytest.mark.parametrize("alpha", [1.0, 0.3])
def test_epsilon_one_gives_dirichlet_distributed_root_priors(alpha):
    """Uniform evaluator, one tick: every game's root has just been expanded with its noised priors."""
    _need_gpu()
    n = 4096
    e, io, L = _engine(n, n, 8, alpha=alpha, eps=1.0)
    e.set_requests(list(range(n)), [0] * n, [0] * n)
    e.eval_builtin(L.EVAL_UNIFORM)
    e.step()
    pri = np.zeros((n, 7), np.float32)
    for s in range(n):
        t = e.dump_tree(s)
        assert t[0] == 2  # expanded root
        pri[s] = t[4:4 + 35].reshape(7, 5)[:, 4].view(np.float32)
    e.close()
    assert (pri >= 0).all() and np.allclose(pri.sum(1), 1.0, atol=2e-6)
    k = 7
    assert np.allclose(pri.mean(0), 1.0 / k, atol=0.01)
    var = (1.0 / k) * (1.0 - 1.0 / k) / (k * alpha + 1.0)
    assert np.allclose(pri.var(0), var, rtol=0.12), (pri.var(0), var)
    assert len({tuple(r) for r in pri.view(np.uint32).tolist()}) == n  # every game draws its own noise


def test_records_are_reproducible_and_independent_of_scheduling():
    _need_gpu()
    from c4a0_b200 import _lib as L

    a = _play(64, 64, 40, alpha=0.5, eps=0.25)
    assert _same(a, _play(64, 64, 40, alpha=0.5, eps=0.25))  # run to run
    assert _same(a, _play(24, 64, 40, alpha=0.5, eps=0.25))  # fewer slots: games are seated later, in other slots
    assert _same(a, _play(64, 64, 40, flags=L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, spec_rows=64, alpha=0.5, eps=0.25))
    assert not _same(a, _play(64, 64, 40))  # and it does change the games
    # the noise belongs to the game id, not to the request index
    ids = [1000 + 3 * i for i in range(64)]
    b = _play(64, 64, 40, alpha=0.5, eps=0.25, ids=ids)
    c = _play(64, 64, 40, alpha=0.5, eps=0.25, ids=ids[::-1])
    for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        assert np.array_equal(getattr(b, f).view(np.uint8), np.ascontiguousarray(getattr(c, f)[::-1]).view(np.uint8))


def _parse(words, idx):
    """Pre-order dump (c4a0_engine_dump_tree): 7 records of 5 words, each expanded child followed by its own node."""
    recs, kids = [], []
    for _ in range(7):
        rec = words[idx:idx + 5]
        idx += 5
        recs.append(rec)
        if rec[0] == 2:
            sub, idx = _parse(words, idx)
            kids.append(sub)
    return (np.array(recs), kids), idx


def test_only_roots_are_noised():
    """With the uniform evaluator every expanded node has priors 1/legal: after some ticks, below the root they must
    still be exactly that, while the root's (epsilon = 1) are the Dirichlet draw; after a move the new root's
    priors are noised as well (subtree reuse: the re-root path)."""
    _need_gpu()
    e, io, L = _engine(8, 8, 24, alpha=1.0, eps=1.0)
    e.set_requests(list(range(8)), [0] * 8, [0] * 8)
    roots_after_move = deeper = 0

    def check_plain(node):
        nonlocal deeper
        recs, kids = node
        legal = recs[:, 0] != 0
        if legal.any():
            want = np.full(int(legal.sum()), np.float32(1.0) / np.float32(legal.sum()), np.float32)
            assert np.array_equal(recs[legal, 4].view(np.float32), want), "a non-root node was noised"
        deeper += 1
        for k in kids:
            check_plain(k)

    for tick in range(80):
        e.eval_builtin(L.EVAL_UNIFORM)
        e.step()
        for s in range(8):
            info = e.slot_info(s)
            t = e.dump_tree(s)
            if info.state == 0 or t[0] != 2:
                continue
            (recs, kids), end = _parse(t, 4)
            assert end == len(t)
            legal = recs[:, 0] != 0
            rp = recs[legal, 4].view(np.float32)
            assert abs(float(rp.sum()) - 1.0) < 2e-6
            if legal.sum() > 1:
                assert not np.allclose(rp, 1.0 / legal.sum(), atol=1e-4)  # noised (equal only with probability ~0)
            for k in kids:
                check_plain(k)
            roots_after_move += int(info.n_moves > 0)
    e.close()
    assert roots_after_move > 0 and deeper > 100
