"""ctypes binding of include/c4a0_engine.h (libc4a0_engine.so).

The shared library is the product; this module only declares its C signatures.  There is no
Python/CPU fallback: if the library is missing (and cannot be built because nvcc is absent) the
import of anything that computes fails with an ImportError.
"""

from __future__ import annotations

import ctypes as C
import os

from . import build as _build

N_ROWS, N_COLS, BUF_N_CHANNELS, PLANE_LEN, MAX_SAMPLES = 6, 7, 2, 84, 43
PLANES_F32, PLANES_BF16 = 0, 1
EVAL_UNIFORM, EVAL_HASH, EVAL_HASH_FLAT = 0, 1, 2
ROW_IDLE, ROW_WAIT_NN, ROW_CONTINUE, ROW_NEED_MOVE = 0, 1, 2, 3
MATH_LOGF, MATH_EXPF = 0, 1
E_INVALID, E_CUDA, E_NOMEM, E_ENGINE = -1, -2, -3, -4
FLAG_NO_DEDUP = 1
FLAG_EVAL_CACHE = 2
FLAG_SPECULATE = 4


class Config(C.Structure):
    _fields_ = [
        ("n_slots", C.c_uint32),
        ("max_requests", C.c_uint32),
        ("n_mcts_iterations", C.c_uint32),
        ("c_exploration", C.c_float),
        ("c_ply_penalty", C.c_float),
        ("plane_dtype", C.c_uint32),
        ("max_inline_sims", C.c_uint32),
        ("device", C.c_int32),
        ("plane_stride", C.c_uint32),
        ("flags", C.c_uint32),
        ("arena_blocks", C.c_uint32),
        ("eval_cache_entries", C.c_uint32),
        ("spec_rows", C.c_uint32),
        ("dirichlet_alpha", C.c_float),
        ("dirichlet_epsilon", C.c_float),
    ]


class Progress(C.Structure):
    _fields_ = [
        ("n_requests", C.c_uint32),
        ("n_started", C.c_uint32),
        ("n_finished", C.c_uint32),
        ("n_running", C.c_uint32),
        ("n_movers", C.c_uint32),
        ("n_rows", C.c_uint32),
        ("error", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("sims", C.c_uint64),
        ("nn_evals", C.c_uint64),
        ("leaf_requests", C.c_uint64),
        ("terminal_leaf_sims", C.c_uint64),
        ("skipped_root_sims", C.c_uint64),
        ("moves", C.c_uint64),
        ("samples", C.c_uint64),
        ("select_depth_sum", C.c_uint64),
        ("expansions", C.c_uint64),
        ("steps", C.c_uint64),
        ("compacted_blocks", C.c_uint64),
        ("compactions", C.c_uint64),
        ("cache_hits", C.c_uint64),
        ("cache_inserts", C.c_uint64),
        ("spec_rows", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SlotInfo(C.Structure):
    _fields_ = [
        ("state", C.c_uint32),
        ("request", C.c_uint32),
        ("n_moves", C.c_uint32),
        ("root_visits", C.c_uint32),
        ("root_mask", C.c_uint64),
        ("root_value", C.c_uint64),
        ("root_q_sum_penalty", C.c_float),
        ("root_q_sum_no_penalty", C.c_float),
        ("n_blocks", C.c_uint32),
        ("nn_row", C.c_uint32),
    ]


class NNGraph(C.Structure):
    _fields_ = [("rows", C.c_uint32), ("graph_exec", C.c_void_p)]


class RunReport(C.Structure):
    _fields_ = [
        ("ticks", C.c_uint64),
        ("nn_launches", C.c_uint64),
        ("nn_rows_launched", C.c_uint64),
        ("nn_relaunches", C.c_uint64),
        ("tail_launches", C.c_uint64),
        ("device_ms", C.c_double),
        ("wall_ms", C.c_double),
        ("host_wait_ms", C.c_double),
        ("host_launch_ms", C.c_double),
        ("kernel_samples", C.c_uint32),
        ("reserved", C.c_uint32),
        ("k_step_ms_sum", C.c_double),
        ("nn_ms_sum", C.c_double),
        ("bucket_launches", C.c_uint64 * 32),
    ]

    def as_dict(self) -> dict:
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["bucket_launches"] = list(self.bucket_launches)
        return d


NET_MAX_LAYERS, NET_MAX_BUFFERS = 12, 8
NET_TILE_M, NET_TILE_N, NET_TILE_K, NET_HEAD_N, NET_PAD_N = 128, 192, 64, 16, 1344
NET_HIDDEN, NET_POLICY, NET_VALUE = 0, 1, 2


class NetLayer(C.Structure):  # include/c4a0_net.h: c4a0_net_layer
    _fields_ = [
        ("weight_dev", C.c_void_p),
        ("bias_dev", C.c_void_p),
        ("n_pad", C.c_uint32),
        ("k_pad", C.c_uint32),
        ("in_buffer", C.c_uint32),
        ("in_col0", C.c_uint32),
        ("out_buffer", C.c_uint32),
        ("out_col0", C.c_uint32),
        ("dep", C.c_int32),
        ("kind", C.c_uint32),
    ]


class NetSpec(C.Structure):  # c4a0_net_spec
    _fields_ = [
        ("device", C.c_int32),
        ("max_rows", C.c_uint32),
        ("n_buffers", C.c_uint32),
        ("buffer_cols", C.c_uint32 * NET_MAX_BUFFERS),
        ("planes_buffer", C.c_uint32),
        ("planes_col0", C.c_uint32),
        ("n_layers", C.c_uint32),
        ("layers", NetLayer * NET_MAX_LAYERS),
    ]


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"c4a0 engine error {code}: {msg}")
        self.code = code


_P = C.c_void_p
# name -> (restype, argtypes); every symbol include/c4a0_engine.h declares
SIGNATURES = {
    "c4a0_last_error": (C.c_char_p, []),
    "c4a0_abi_version": (C.c_int, []),
    "c4a0_engine_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "c4a0_engine_destroy": (None, [_P]),
    "c4a0_engine_io_rows": (C.c_uint32, [_P]),
    "c4a0_engine_device_bytes": (C.c_size_t, [_P]),
    "c4a0_engine_bind_io": (C.c_int, [_P, _P, _P, _P, _P]),
    "c4a0_engine_set_requests": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P]),
    "c4a0_engine_step": (C.c_int, [_P, _P]),
    "c4a0_engine_step_timed": (C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "c4a0_engine_debug_phases": (C.c_int, [_P, _P, _P]),
    "c4a0_engine_eval_builtin": (C.c_int, [_P, C.c_int, _P]),
    "c4a0_engine_poll": (C.c_int, [_P, C.POINTER(Progress), _P]),
    "c4a0_engine_stats": (C.c_int, [_P, C.POINTER(Stats), _P]),
    "c4a0_engine_fetch_rows": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "c4a0_engine_rows_dev": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "c4a0_engine_rows_count_dev": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "c4a0_engine_run": (C.c_int, [_P, C.c_uint32, _P, _P, _P, C.c_uint64, C.c_uint32, C.POINTER(RunReport)]),
    "c4a0_engine_run_net": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, C.c_uint32, C.POINTER(RunReport)]),
    "c4a0_engine_fetch_results": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P, _P]),
    "c4a0_engine_export_samples": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.c_int, _P, _P, _P, _P, _P]),
    "c4a0_engine_results_dev": (C.c_int, [_P] + [C.POINTER(_P)] * 6),
    "c4a0_engine_slot_info": (C.c_int, [_P, C.c_uint32, C.POINTER(SlotInfo), _P]),
    "c4a0_engine_dump_tree": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t), _P]),
    "c4a0_rules_batch": (C.c_int, [C.c_int, _P, _P, C.c_size_t, C.c_float] + [_P] * 10),
    "c4a0_math_batch": (C.c_int, [C.c_int, C.c_int, _P, _P, C.c_size_t]),
    "c4a0_softmax_batch": (C.c_int, [C.c_int, _P, _P, _P, C.c_size_t]),
    "c4a0_heads": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P, C.c_uint32, _P, _P, _P, _P]),
    "c4a0_head_epilogue": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "c4a0_sample_batch": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_size_t]),
    "c4a0_host_logf": (None, [_P, _P, C.c_size_t]),
    "c4a0_host_expf": (None, [_P, _P, C.c_size_t]),
    "c4a0_host_sample": (C.c_int, [_P, C.c_float, C.c_uint64, _P]),
    "c4a0_host_terminal_state": (C.c_int, [C.c_uint64, C.c_uint64]),
    "c4a0_host_make_move": (None, [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "c4a0_host_pos_key": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "c4a0_host_flip_h": (None, [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "c4a0_host_shuffle": (None, [C.c_uint64, _P, C.c_size_t]),
    "c4a0_host_seed_to_key": (None, [C.c_uint64, _P]),
    "c4a0_host_stdrng_words": (None, [_P, _P, C.c_size_t]),
    "c4a0_results_to_cbor": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "c4a0_results_from_cbor": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint32), _P, _P, _P, _P, _P, _P, _P]),
    # include/c4a0_net.h
    "c4a0_net_create": (C.c_int, [C.POINTER(NetSpec), C.POINTER(_P)]),
    "c4a0_net_destroy": (None, [_P]),
    "c4a0_net_device_bytes": (C.c_size_t, [_P]),
    "c4a0_net_buffer": (C.c_int, [_P, C.c_uint32, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "c4a0_net_bind_outputs": (C.c_int, [_P, _P, _P, _P]),
    "c4a0_net_bind_row_count": (C.c_int, [_P, _P, _P]),
    "c4a0_net_forward": (C.c_int, [_P, C.c_uint32, _P]),
    "c4a0_net_forward_ex": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32]),
    "c4a0_net_debug_trace": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, C.c_size_t]),
    "c4a0_net_forward_timed": (C.c_int, [_P, C.c_uint32, _P, C.POINTER(C.c_float)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (building first if the .so is absent) and type the C-ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        try:
            _build.build_engine()
        except Exception as exc:  # no nvcc and no prebuilt library: nothing can run
            raise ImportError(f"libc4a0_engine.so is missing and could not be built: {exc}") from exc
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export the header
        fn.restype = res
        fn.argtypes = args
    if L.c4a0_abi_version() != 3:
        raise ImportError("libc4a0_engine.so ABI version mismatch; rebuild with c4a0_b200/build.py")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(rc, (lib().c4a0_last_error() or b"").decode("utf-8", "replace"))


def ptr(arr) -> C.c_void_p:
    """void* of a numpy array (host) — None passes NULL."""
    if arr is None:
        return None
    return C.c_void_p(arr.ctypes.data)
