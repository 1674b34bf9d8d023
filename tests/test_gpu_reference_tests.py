"""The reference's own boundary test, byte for byte, against this repo's `c4a0_rust`.

`tests/golden/pybridge_test.py` is a verbatim copy of /root/reference/tests/c4a0_tests/pybridge_test.py
(sha256 pinned below; where the reference tree is present the copy is also compared with it).  It needs
only `c4a0_rust` + numpy and drives `play_games` through the numpy-callback contract
(rust/src/pybridge.rs:161-199), so it runs unmodified.  The other reference test modules import
`c4a0.nn` / `c4a0.training`, i.e. pytorch_lightning + torchmetrics, which this image does not have.
"""

import hashlib
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
COPY = os.path.join(HERE, "golden", "pybridge_test.py")
SHA256 = "030b697a32f7777dd754ae01bcd4313ad9bb2bc8149ec72df408646a2a2663ce"
REFERENCE = "/root/reference/tests/c4a0_tests/pybridge_test.py"


def test_copy_is_the_reference_file():
    data = open(COPY, "rb").read()
    assert hashlib.sha256(data).hexdigest() == SHA256
    if os.path.exists(REFERENCE):  # absent on the GPU box
        assert open(REFERENCE, "rb").read() == data


@pytest.mark.gpu
def test_reference_pybridge_test_runs_unmodified():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    spec = importlib.util.spec_from_file_location("reference_pybridge_test", COPY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tests = [getattr(mod, n) for n in dir(mod) if n.startswith("test_") and callable(getattr(mod, n))]
    assert tests, "the reference module defines tests"
    for t in tests:
        t()
