"""Python handle over the C-ABI engine (include/c4a0_engine.h).

Host-side mirror of what `self_play::self_play` (rust/src/self_play.rs:39-129) owns in the
reference: the set of games, their trees and the finished samples — here all of it lives in HBM
behind a `c4a0_engine*`.  Device buffers are passed as raw pointers (`tensor.data_ptr()`), streams as
`torch.cuda.Stream.cuda_stream`; torch itself is not needed by this module.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


@dataclass
class GameSamples:
    """Finished games in request order as struct-of-arrays (types.rs:104-110 `Sample` fields)."""

    n_samples: np.ndarray  # [G] uint32, 0 = game not finished
    mask: np.ndarray  # [G,43] uint64
    value: np.ndarray  # [G,43] uint64
    policy: np.ndarray  # [G,43,7] float32
    q_penalty: np.ndarray  # [G,43] float32
    q_no_penalty: np.ndarray  # [G,43] float32


class Engine:
    def __init__(
        self,
        n_slots: int,
        max_requests: int,
        n_mcts_iterations: int,
        c_exploration: float,
        c_ply_penalty: float,
        plane_dtype: int = L.PLANES_F32,
        max_inline_sims: int = 0,
        device: int = 0,
        plane_stride: int = 0,
        flags: int = 0,
        arena_blocks: int = 0,
        eval_cache_entries: int = 0,
        spec_rows: int = 0,
        dirichlet_alpha: float = 0.0,
        dirichlet_epsilon: float = 0.0,
    ):
        self._lib = L.lib()
        self._h = C.c_void_p()
        self.cfg = L.Config(
            n_slots, max_requests, n_mcts_iterations, c_exploration, c_ply_penalty, plane_dtype, max_inline_sims, device,
            plane_stride, flags, arena_blocks, eval_cache_entries, spec_rows, dirichlet_alpha, dirichlet_epsilon,
        )
        L.check(self._lib.c4a0_engine_create(C.byref(self.cfg), C.byref(self._h)))
        self.n_slots = n_slots
        self.io_rows = int(self._lib.c4a0_engine_io_rows(self._h))  # rows the NN I/O buffers must have
        self.n_requests = 0

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.c4a0_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def device_bytes(self) -> int:
        return int(self._lib.c4a0_engine_device_bytes(self._h))

    def bind_io(self, planes_ptr: int, logits_ptr: int, qp_ptr: int, qn_ptr: int) -> None:
        L.check(self._lib.c4a0_engine_bind_io(self._h, planes_ptr, logits_ptr, qp_ptr, qn_ptr))

    def set_requests(self, game_id, player0_id, player1_id, stream: int = 0) -> None:
        g = np.ascontiguousarray(game_id, dtype=np.uint64)
        a = np.ascontiguousarray(player0_id, dtype=np.uint64)
        b = np.ascontiguousarray(player1_id, dtype=np.uint64)
        if not (g.shape == a.shape == b.shape and g.ndim == 1):
            raise ValueError("game_id / player0_id / player1_id must be 1-d arrays of equal length")
        L.check(self._lib.c4a0_engine_set_requests(self._h, L.ptr(g), L.ptr(a), L.ptr(b), len(g), stream))
        self.n_requests = len(g)

    def step(self, stream: int = 0) -> None:
        L.check(self._lib.c4a0_engine_step(self._h, stream))

    def step_timed(self, stream: int = 0) -> Tuple[float, float]:
        """step() bracketed by CUDA events: (ms of k_step, ms of k_tail)."""
        a, b = C.c_float(), C.c_float()
        L.check(self._lib.c4a0_engine_step_timed(self._h, stream, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_phases(self, stream: int = 0) -> np.ndarray:
        """One step() with per-game cycle counters: [n_slots, 8] uint32 (see the header)."""
        out = np.zeros((self.n_slots, 8), np.uint32)
        L.check(self._lib.c4a0_engine_debug_phases(self._h, stream, L.ptr(out)))
        return out

    def eval_builtin(self, kind: int, stream: int = 0) -> None:
        L.check(self._lib.c4a0_engine_eval_builtin(self._h, kind, stream))

    def poll(self, stream: int = 0) -> L.Progress:
        p = L.Progress()
        L.check(self._lib.c4a0_engine_poll(self._h, C.byref(p), stream))
        return p

    def stats(self, stream: int = 0) -> dict:
        s = L.Stats()
        L.check(self._lib.c4a0_engine_stats(self._h, C.byref(s), stream))
        return s.as_dict()

    def fetch_rows(self, stream: int = 0):
        """(n_rows, leaf_mask[n_rows], leaf_value[n_rows], model_id[n_rows]) of the live rows."""
        S = self.io_rows
        n = C.c_uint32(0)
        mask = np.empty(S, np.uint64)
        value = np.empty(S, np.uint64)
        model = np.empty(S, np.uint64)
        L.check(self._lib.c4a0_engine_fetch_rows(self._h, C.byref(n), L.ptr(mask), L.ptr(value), L.ptr(model), stream))
        k = int(n.value)
        return k, mask[:k], value[:k], model[:k]

    def fetch_results(self, first: int = 0, n: Optional[int] = None, stream: int = 0) -> GameSamples:
        if n is None:
            n = self.n_requests - first
        M = L.MAX_SAMPLES
        out = GameSamples(
            np.zeros(n, np.uint32),
            np.zeros((n, M), np.uint64),
            np.zeros((n, M), np.uint64),
            np.zeros((n, M, 7), np.float32),
            np.zeros((n, M), np.float32),
            np.zeros((n, M), np.float32),
        )
        if n:
            L.check(
                self._lib.c4a0_engine_fetch_results(
                    self._h, first, n, L.ptr(out.n_samples), L.ptr(out.mask), L.ptr(out.value), L.ptr(out.policy),
                    L.ptr(out.q_penalty), L.ptr(out.q_no_penalty), stream,
                )
            )
        return out

    def results_dev(self) -> Tuple[int, int, int, int, int, int]:
        """Device pointers (n_samples, mask, value, policy, q_penalty, q_no_penalty)."""
        ps = [C.c_void_p() for _ in range(6)]
        L.check(self._lib.c4a0_engine_results_dev(self._h, *[C.byref(p) for p in ps]))
        return tuple(int(p.value) for p in ps)

    def export_samples(self, first: int, n: int, offsets_ptr: int, total: int, flip: bool, pos_ptr: int,
                       policy_ptr: int, qp_ptr: int, qn_ptr: int, stream: int = 0) -> None:
        """Device-side training tensors; see c4a0_engine_export_samples in the header."""
        L.check(self._lib.c4a0_engine_export_samples(self._h, first, n, offsets_ptr, total, int(flip), pos_ptr,
                                                     policy_ptr, qp_ptr, qn_ptr, stream))

    def rows_dev(self) -> Tuple[int, int]:
        """Device pointers (row_slot u32[io_rows], row_model u64[io_rows])."""
        a, b = C.c_void_p(), C.c_void_p()
        L.check(self._lib.c4a0_engine_rows_dev(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def rows_count_dev(self) -> Tuple[int, int]:
        """Device addresses (closed, open) of the batch's row count; see c4a0_engine_rows_count_dev."""
        a, b = C.c_void_p(), C.c_void_p()
        L.check(self._lib.c4a0_engine_rows_count_dev(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def slot_info(self, slot: int, stream: int = 0) -> L.SlotInfo:
        info = L.SlotInfo()
        L.check(self._lib.c4a0_engine_slot_info(self._h, slot, C.byref(info), stream))
        return info

    def dump_tree(self, slot: int, stream: int = 0) -> np.ndarray:
        need = C.c_size_t(0)
        L.check(self._lib.c4a0_engine_dump_tree(self._h, slot, None, 0, C.byref(need), stream))
        buf = np.zeros(need.value, np.uint32)
        L.check(self._lib.c4a0_engine_dump_tree(self._h, slot, L.ptr(buf), buf.size, C.byref(need), stream))
        return buf


def run_engines(engines: Sequence["Engine"], graphs: Sequence[Sequence[Tuple[int, int]]], streams: Sequence[int],
                max_ticks: int = 0, time_kernels_every: int = 0) -> dict:
    """c4a0_engine_run(): graphs[i] = [(rows, cudaGraphExec_t as int), ...] sorted by rows."""
    n = len(engines)
    handles = (C.c_void_p * n)(*[e._h for e in engines])
    arrays = [(L.NNGraph * len(g))(*[L.NNGraph(r, h) for r, h in g]) for g in graphs]
    gptr = (C.c_void_p * n)(*[C.cast(a, C.c_void_p) for a in arrays])
    counts = (C.c_uint32 * n)(*[len(g) for g in graphs])
    st = (C.c_void_p * n)(*streams)
    rep = L.RunReport()
    L.check(L.lib().c4a0_engine_run(handles, n, gptr, counts, st, max_ticks, time_kernels_every, C.byref(rep)))
    return rep.as_dict()


def run_engines_net(engines: Sequence["Engine"], nets: Sequence[int], streams: Sequence[int], max_ticks: int = 0,
                    time_kernels_every: int = 0) -> dict:
    """c4a0_engine_run_net(): nets[i] = c4a0_net handle (as int) evaluating engine i; both kernels of a tick are
    launched directly by the C++ loop, programmatically dependent on each other."""
    n = len(engines)
    handles = (C.c_void_p * n)(*[e._h for e in engines])
    nh = (C.c_void_p * n)(*nets)
    st = (C.c_void_p * n)(*streams)
    rep = L.RunReport()
    L.check(L.lib().c4a0_engine_run_net(handles, n, nh, st, max_ticks, time_kernels_every, C.byref(rep)))
    return rep.as_dict()


# ------------------------------------------------------------------------------------------------
# stand-alone batch kernels (host arrays in, host arrays out)
# ------------------------------------------------------------------------------------------------
def rules_batch(mask: Sequence[int], value: Sequence[int], c_ply_penalty: float = 0.01, device: int = 0) -> dict:
    m = np.ascontiguousarray(mask, dtype=np.uint64)
    v = np.ascontiguousarray(value, dtype=np.uint64)
    n = len(m)
    out = dict(
        terminal=np.zeros(n, np.int32),
        legal=np.zeros(n, np.uint32),
        ply=np.zeros(n, np.int32),
        q_penalty=np.zeros(n, np.float32),
        q_no_penalty=np.zeros(n, np.float32),
        child_mask=np.zeros((n, 7), np.uint64),
        child_value=np.zeros((n, 7), np.uint64),
        planes=np.zeros((n, 2, 6, 7), np.float32),
        flip_mask=np.zeros(n, np.uint64),
        flip_value=np.zeros(n, np.uint64),
    )
    L.check(
        L.lib().c4a0_rules_batch(
            device, L.ptr(m), L.ptr(v), n, c_ply_penalty, L.ptr(out["terminal"]), L.ptr(out["legal"]), L.ptr(out["ply"]),
            L.ptr(out["q_penalty"]), L.ptr(out["q_no_penalty"]), L.ptr(out["child_mask"]), L.ptr(out["child_value"]),
            L.ptr(out["planes"]), L.ptr(out["flip_mask"]), L.ptr(out["flip_value"]),
        )
    )
    return out


def math_batch(op: int, x: np.ndarray, device: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    L.check(L.lib().c4a0_math_batch(device, op, L.ptr(x), L.ptr(out), x.size))
    return out


def softmax_batch(logits: np.ndarray, legal: np.ndarray, device: int = 0) -> np.ndarray:
    lg = np.ascontiguousarray(logits, dtype=np.float32).reshape(-1, 7)
    le = np.ascontiguousarray(legal, dtype=np.uint32)
    out = np.empty_like(lg)
    L.check(L.lib().c4a0_softmax_batch(device, L.ptr(lg), L.ptr(le), L.ptr(out), len(lg)))
    return out


def sample_batch(policy: np.ndarray, temperature: np.ndarray, seed: np.ndarray, device: int = 0):
    p = np.ascontiguousarray(policy, dtype=np.float32).reshape(-1, 7)
    t = np.ascontiguousarray(temperature, dtype=np.float32)
    s = np.ascontiguousarray(seed, dtype=np.uint64)
    tempered = np.empty_like(p)
    col = np.empty(len(p), np.int32)
    L.check(L.lib().c4a0_sample_batch(device, L.ptr(p), L.ptr(t), L.ptr(s), L.ptr(tempered), L.ptr(col), len(p)))
    return tempered, col


def host_logf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    L.lib().c4a0_host_logf(L.ptr(x), L.ptr(out), x.size)
    return out


def host_expf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    L.lib().c4a0_host_expf(L.ptr(x), L.ptr(out), x.size)
    return out


def host_sample(policy: Sequence[float], temperature: float, seed: int):
    p = np.ascontiguousarray(policy, dtype=np.float32)
    t = np.empty(7, np.float32)
    col = L.lib().c4a0_host_sample(L.ptr(p), temperature, seed, L.ptr(t))
    return t, int(col)


def host_flip_h(mask: int, value: int) -> Tuple[int, int]:
    a, b = C.c_uint64(), C.c_uint64()
    L.lib().c4a0_host_flip_h(mask, value, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def host_shuffle(seed: int, n: int) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint32)
    L.lib().c4a0_host_shuffle(seed & 0xFFFFFFFFFFFFFFFF, L.ptr(idx), n)
    return idx


def host_terminal_state(mask: int, value: int) -> int:
    return int(L.lib().c4a0_host_terminal_state(mask, value))
