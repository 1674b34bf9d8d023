"""Generates tests/golden/selfplay_games.npz: complete game records of the CPU oracle (which is pinned to
the reference's known-answer tests, tests/test_oracle_*.py) for fixed requests and both synthetic
evaluators.  The GPU engine must reproduce these files bit for bit (tests/test_gpu_golden.py), and so
must the oracle itself (tests/test_oracle_golden.py), so a change of either shows up against a
committed artefact, not only against the other."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = {
    # name: (evaluator, n_games, n_iter, c_exploration, c_ply_penalty, game ids, player ids)
    "uniform": ("uniform", 12, 50, 4.0, 0.01, [3 * i + 1 for i in range(12)], (0, 0)),
    "hash": ("hash", 16, 64, 6.6, 0.01, [1000 + 37 * i for i in range(16)], (0, 0)),
    "hash_two_models": ("hash", 10, 40, 6.6, 0.01, [2**40 + 5 * i for i in range(10)], (7, 9)),
}


def records_to_arrays(samples):
    n = len(samples)
    out = dict(
        n_samples=np.zeros(n, np.uint32), mask=np.zeros((n, 43), np.uint64), value=np.zeros((n, 43), np.uint64),
        policy=np.zeros((n, 43, 7), np.float32), q_penalty=np.zeros((n, 43), np.float32),
        q_no_penalty=np.zeros((n, 43), np.float32),
    )
    for i, game in enumerate(samples):
        out["n_samples"][i] = len(game)
        for k, s in enumerate(game):
            out["mask"][i, k], out["value"][i, k] = s.pos.mask, s.pos.value
            out["policy"][i, k] = np.array(list(s.policy), np.float32)
            out["q_penalty"][i, k], out["q_no_penalty"][i, k] = s.q_penalty, s.q_no_penalty
    return out


def run_case(name):
    ev, n, n_iter, c_expl, c_pen, ids, (p0, p1) = CASES[name]
    reqs = [(g, p0, p1) for g in ids]
    o = oracle.self_play(reqs, n, n_iter, c_expl, c_pen, evaluator=ev)
    return records_to_arrays(o.samples)


if __name__ == "__main__":
    blob = {}
    for name in CASES:
        for k, v in run_case(name).items():
            blob[f"{name}/{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "selfplay_games.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes;", {n: int(blob[f"{n}/n_samples"].sum()) for n in CASES}, "samples")
