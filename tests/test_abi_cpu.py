"""The C-ABI library loads on a CPU-only box and exports every symbol include/c4a0_engine.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("c4a0_engine.h", "c4a0_net.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(c4a0_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported_and_bound():
    from c4a0_b200 import _lib as L

    lib = L.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.c4a0_abi_version() == 3


def test_struct_layouts_match_the_compiled_headers(tmp_path):
    """tests/abi_check.c includes include/c4a0_engine.h and include/c4a0_net.h, is compiled with gcc (C99, no
    CUDA) and prints sizeof / offsetof of every struct field; the ctypes mirrors must agree field by field."""
    import shutil
    import subprocess

    from c4a0_b200 import _lib as L

    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    exe = tmp_path / "abi_check"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_check.c"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {ln.split()[0]: tuple(int(x) for x in ln.split()[1:]) for ln in out if ln.strip()}
    mirrors = {"c4a0_config": L.Config, "c4a0_progress": L.Progress, "c4a0_stats": L.Stats, "c4a0_nn_graph": L.NNGraph,
               "c4a0_run_report": L.RunReport, "c4a0_slot_info": L.SlotInfo, "c4a0_net_layer": L.NetLayer,
               "c4a0_net_spec": L.NetSpec}
    for cname, st in mirrors.items():
        assert got[cname] == (C.sizeof(st),), cname
        for fname, _ in st._fields_:
            f = getattr(st, fname)
            assert got[f"{cname}.{fname}"] == (f.offset, f.size), f"{cname}.{fname}"
        # the header has no field the mirror lacks: the fields tile the struct up to alignment padding
        assert sum(1 for k in got if k.startswith(cname + ".")) == len(st._fields_)
    assert got["C4A0_ABI_VERSION"] == (L.lib().c4a0_abi_version(),)
    assert got["C4A0_MAX_SAMPLES"] == (L.MAX_SAMPLES,) and got["C4A0_NET_MAX_LAYERS"] == (L.NET_MAX_LAYERS,)
    assert got["C4A0_NET_PAD_N"] == (L.NET_PAD_N,)


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from c4a0_b200 import _lib as L
    from c4a0_b200 import engine as E

    with pytest.raises(L.EngineError) as ei:
        E.Engine(4, 4, 10, 1.0, 0.01)
    assert ei.value.code == L.E_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(L.EngineError):
        E.rules_batch([0], [0])
    with pytest.raises(L.EngineError):
        E.math_batch(L.MATH_LOGF, np.ones(4, np.float32))
    import c4a0_rust as R

    with pytest.raises(RuntimeError):
        R.play_games([R.GameMetadata(0, 0, 0)], 4, 2, 1.0, 0.01, lambda m, p: None)


def test_product_does_not_import_the_oracle():
    bad = []
    for pkg in ("c4a0_b200", "c4a0_rust"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M) or "c4a0_oracle" in txt or "c4o_" in txt:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
