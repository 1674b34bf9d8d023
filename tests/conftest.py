import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(autouse=True)
def _seed():
    # tests/c4a0_tests/conftest.py:7-11 of the reference seeds everything with 1337
    random.seed(1337)
    np.random.seed(1337)
    try:
        import torch

        torch.manual_seed(1337)
    except Exception:
        pass


# tests/golden/ holds fixtures (and the reference's own pybridge_test.py, byte for byte); they are run by
# the tests that name them, not collected on their own
collect_ignore_glob = ["golden/*"]
