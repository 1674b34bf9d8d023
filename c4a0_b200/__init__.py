"""c4a0_b200 — B200-native self-play engine behind the reference's `c4a0_rust.play_games`.

csrc/      hand-written sm_100a CUDA kernels + the C-ABI (include/c4a0_engine.h)
_lib.py    ctypes signatures of the C-ABI
engine.py  host handle over the engine
"""

__all__ = ["build", "engine"]
