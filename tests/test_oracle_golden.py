"""The oracle reproduces the committed golden game records (tests/golden/selfplay_games.npz, written by
tools/gen_golden_games.py): a change of the oracle shows up against a committed artefact."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden_games as G  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "selfplay_games.npz")


def test_oracle_reproduces_golden_games():
    blob = np.load(GOLDEN)
    for name in G.CASES:
        got = G.run_case(name)
        for k, v in got.items():
            exp = blob[f"{name}/{k}"]
            assert v.dtype == exp.dtype and v.shape == exp.shape, (name, k)
            assert v.tobytes() == exp.tobytes(), (name, k)  # bit patterns, floats included


def test_golden_games_are_well_formed():
    blob = np.load(GOLDEN)
    for name in G.CASES:
        n = blob[f"{name}/n_samples"]
        assert (n >= 8).all() and (n <= 43).all()  # a game lasts at least 7 plies + the terminal sample
        pol = blob[f"{name}/policy"]
        for i in range(len(n)):
            p = pol[i, : n[i]]
            assert np.allclose(p.sum(1), 1.0, atol=1e-5)
            assert not pol[i, n[i] :].any()
