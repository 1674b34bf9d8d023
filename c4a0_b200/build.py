"""Builds libc4a0_engine.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

The library has no torch / Python dependency, so this is one nvcc invocation.  -fmad=false is part
of the numerics contract: the reference never contracts f32 a*b+c; the double
FMAs inside c4_logf/c4_expf are explicit fma() calls and are unaffected.
"""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libc4a0_engine.so")

SOURCES = ["engine.cu", "batch_ops.cu", "net.cu", "cbor.cu"]
HEADERS = ["c4_rules.cuh", "c4_math.cuh", "c4_rng.cuh", "common.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libc4a0_engine.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", h) for h in ("c4a0_engine.h", "c4a0_net.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_engine(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv))
