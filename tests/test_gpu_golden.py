"""The GPU engine reproduces the committed golden game records (tests/golden/selfplay_games.npz) bit for
bit — plain, with the evaluation cache, and with speculative rows — without the oracle in the loop."""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden_games import CASES  # noqa: E402  (the requests of every case; nothing of the oracle is called)

GOLDEN = os.path.join(ROOT, "tests", "golden", "selfplay_games.npz")


@pytest.mark.parametrize("mode", ["plain", "cache", "speculate"])
@pytest.mark.parametrize("name", list(CASES))
def test_engine_reproduces_golden_games(name, mode):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from c4a0_b200 import _lib as L
    from c4a0_b200.engine import Engine

    ev, n, n_iter, c_expl, c_pen, ids, (p0, p1) = CASES[name]
    flags = {"plain": 0, "cache": L.FLAG_EVAL_CACHE, "speculate": L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE}[mode]
    n_slots = max(2, n - 3)  # fewer slots than games: some games are seated when others finish
    e = Engine(n_slots, n, n_iter, c_expl, c_pen, flags=flags, spec_rows=16 if mode == "speculate" else 0)
    R = e.io_rows
    io = [torch.zeros(R, 2, 6, 7, device="cuda"), torch.zeros(R, 7, device="cuda"), torch.zeros(R, device="cuda"),
          torch.zeros(R, device="cuda")]
    e.bind_io(*[t.data_ptr() for t in io])
    e.set_requests(ids, [p0] * n, [p1] * n)
    kind = L.EVAL_UNIFORM if ev == "uniform" else L.EVAL_HASH
    for t in range(200000):
        e.eval_builtin(kind)
        e.step()
        if t % 16 == 15 and e.poll().n_finished == n:
            break
    got = e.fetch_results()
    blob = np.load(GOLDEN)
    for k in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty"):
        exp = blob[f"{name}/{k}"]
        g = np.asarray(getattr(got, k))
        assert g.shape == exp.shape, (k, g.shape, exp.shape)
        assert g.astype(exp.dtype).tobytes() == exp.tobytes(), k
    e.close()
