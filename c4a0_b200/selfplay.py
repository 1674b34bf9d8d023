"""Lockstep self-play driver: the engine's tree kernels + the network on CUDA streams.

This is the B200 replacement for the thread/channel machinery of rust/src/self_play.rs:39-246
(`self_play`, `NNThread`).  There the NN thread collects leaf positions from a queue, de-duplicates
them, builds a numpy batch, calls Python, and fans results back through another queue; here a
*tick* is

    network(planes[:B]) -> logits, q_penalty, q_no_penalty
    engine tick         -> expand + backup + move + select + dedup/pack the next planes (k_step)

on the same device buffers, driven by the C++ host loop until every game has finished — no Python in
the loop.  The network is either
  * the library's own kernel (`native_net.NativeEvaluator`, csrc/net.cu: bf16, one launch per tick that
    reads the batch size on the device; `c4a0_engine_run_net` launches both kernels of a tick, chained by
    programmatic dependent launch) — the default for a bf16 `ConnectFourNet`; or
  * a PyTorch callable captured once per batch-size bucket into CUDA graphs (`c4a0_engine_run` picks the
    smallest graph that covers the tick's rows): fp32 networks, several models in one batch, anything else.
With `n_lanes=2` the games are split over two engines on two streams (measured: no gain).

Evaluator contracts:
  * device evaluators (fast path): `NativeEvaluator`, or `fn(planes: cuda Tensor[B, stride]) ->
    (policy[B,7], qp[B], qn[B])` cuda tensors — `DeviceEvaluator.from_model(ConnectFourNet)` picks the form;
  * the reference's numpy callback `cb(model_id, ndarray[B,2,6,7]) -> (policy, qp, qn)`
    (rust/src/pybridge.rs:161-199), served by `play_callback` with host copies every tick.
"""

from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import Engine, GameSamples, run_engines, run_engines_net

# process-wide defaults, overridable by tools (bench.py turns kernel sampling on)
DEFAULTS = {"host_loop": "native", "poll_every": 32, "sample_kernels_every": 0, "n_lanes": 1, "dedup": True,
            "max_inline_sims": 0, "arena_blocks": None,
            # evaluation cache (engine.cu, C4A0_FLAG_EVAL_CACHE) for device evaluators reached through
            # c4a0_rust.play_games; callbacks never get it (they need not be pure functions of the position)
            "eval_cache": True, "eval_cache_entries": 0,
            # with the cache: top small batches up with the children of expanded leaves (C4A0_FLAG_SPECULATE)
            "speculate": True, "spec_rows": 0,
            # False: play_games leaves the samples in the engine's device store (multi-GPU ranks whose samples
            # travel to rank 0 over NCCL, trainers that read export_tensors()) and returns an empty result
            "fetch": True,
            # (alpha, epsilon) of Dirichlet noise on root priors, or None: off, like the reference (c4a0_config)
            "dirichlet": None}

def _buckets():
    """Network batch sizes that get their own CUDA graph: fine steps (a tick launches the smallest
    graph covering its live rows, so coarse steps waste GEMM rows), a few percent apart above 1024."""
    b = [128, 256, 384, 512, 768, 1024]
    b += list(range(1280, 4096 + 1, 256))
    b += list(range(4608, 16384 + 1, 512))
    b += list(range(17408, 32768 + 1, 1024))
    b += list(range(36864, 131072 + 1, 4096))
    return tuple(b)


BUCKETS = _buckets()


class DeviceEvaluator:
    """A network the engine can call without leaving the GPU.

    fn(planes) receives a cuda tensor [B, plane_stride] (plane_stride == 84: viewed as [B,2,6,7])
    of dtype `dtype` and returns (policy [B,7], q_penalty [B], q_no_penalty [B]) cuda tensors.
    """

    def __init__(self, module_or_fn, dtype: torch.dtype = torch.float32, plane_stride: int = 84, plane_offset: int = 0):
        self.fn = module_or_fn
        self.dtype = dtype
        self.plane_stride = plane_stride  # elements per row of the buffer the engine writes planes into
        self.plane_offset = plane_offset  # first plane element of a row (the buffer may hold more)
        if isinstance(module_or_fn, torch.nn.Module) and module_or_fn.training:
            # a module evaluated as it is (no folding): self-play needs inference behaviour (BatchNorm running
            # statistics, no dropout).  FoldedNet / FusedNet / NativeEvaluator never touch the caller's module.
            module_or_fn.eval()

    @classmethod
    def from_model(cls, model, dtype: torch.dtype = torch.bfloat16, fold: bool = True,
                   reuse: Optional["DeviceEvaluator"] = None) -> "DeviceEvaluator":
        """ConnectFourNet -> evaluator.  fold=True uses the GEMM-folded inference form (nn.FoldedNet).
        Pass the previous generation's evaluator as `reuse` to load the new weights into it in place:
        the CUDA graphs captured over it (and the cached session) stay valid."""
        from .native_net import NativeEvaluator
        from .nn import ConnectFourNet, FoldedNet, FusedNet

        if (fold is True and dtype == torch.bfloat16 and NativeEvaluator.supports(model)
                and os.environ.get("C4A0_NATIVE_NET", "1") != "0"):
            # the library's own tcgen05 kernel (csrc/net.cu) evaluates the folded network in one launch
            if isinstance(reuse, NativeEvaluator):
                try:
                    return reuse.refresh(model)
                except ValueError:
                    pass
            return NativeEvaluator(model)
        if fold == "cublas":  # the same folded network through PyTorch / cuBLASLt (comparison, multi-model)
            fold = True
        if fold and isinstance(model, ConnectFourNet):
            if isinstance(reuse, DeviceEvaluator) and isinstance(reuse.fn, FoldedNet) and reuse.dtype == dtype:
                try:
                    reuse.fn.refresh(model)
                    return reuse
                except ValueError:
                    pass
            if fold != "plain" and FusedNet.supports(model):
                net = FusedNet(model, dtype=dtype)
                return cls(net, dtype, net.plane_stride, net.plane_offset)
            return cls(FoldedNet(model, dtype=dtype), dtype, FoldedNet.IN_PAD)
        p = next(model.parameters(), None)
        return cls(model, p.dtype if p is not None else dtype, 84)

    def __call__(self, planes: torch.Tensor):
        if self.plane_stride == 84 and planes.dim() == 2:
            planes = planes.view(-1, 2, 6, 7)
        return self.fn(planes)

    def into(self, planes: torch.Tensor, logits: torch.Tensor, qp: torch.Tensor, qn: torch.Tensor) -> None:
        """Evaluate `planes` and leave the answers in the engine's buffers.  The folded network forms
        write them with one fused output kernel (nn._output_stage); anything else is copied."""
        from .nn import FoldedNet

        rows = planes.shape[0]
        if isinstance(self.fn, FoldedNet):
            self.fn(planes, out=(logits, qp, qn))
            return
        pol, a, b = self(planes)
        logits.copy_(pol.reshape(rows, 7))
        qp.copy_(a.reshape(rows))
        qn.copy_(b.reshape(rows))


class BuiltinEvaluator:
    """The engine's synthetic evaluators (k_eval_builtin: uniform E0, integer-hash E1, flat hash) as a device
    evaluator, so that parity jobs run through the same native host loop, CUDA graphs, evaluation cache and
    speculation as a real network.  A pure function of (model, position); float32 planes (unused)."""

    KINDS = {"uniform": L.EVAL_UNIFORM, "hash": L.EVAL_HASH, "hash_flat": L.EVAL_HASH_FLAT}

    def __init__(self, kind: str = "hash"):
        self.kind = self.KINDS[kind]
        self.dtype = torch.float32
        self.plane_stride = 84
        self.plane_offset = 0

    def __call__(self, planes):  # only reachable through _Lane.evaluate, which launches the kernel itself
        raise TypeError("BuiltinEvaluator is evaluated by the engine (c4a0_engine_eval_builtin)")


class MultiModelEvaluator:
    """Several networks in one batch (tournaments: `GameMetadata.player0_id != player1_id`,
    src/c4a0/tournament.py:112-142).  The reference's NN thread serves one model per callback
    (self_play.rs:211-220); here every model evaluates the tick's live rows and each row keeps the
    answer of the model that has to move in it (`row_model`, mcts.rs:70-76).  All evaluators must
    share dtype and plane layout."""

    def __init__(self, evaluators):
        self.evaluators = {int(k): v for k, v in evaluators.items()}
        if not self.evaluators:
            raise ValueError("no evaluators")
        first = next(iter(self.evaluators.values()))
        self.dtype, self.plane_stride, self.plane_offset = first.dtype, first.plane_stride, getattr(first, "plane_offset", 0)
        for ev in self.evaluators.values():
            if not isinstance(ev, DeviceEvaluator):
                raise TypeError("a MultiModelEvaluator combines DeviceEvaluators (PyTorch networks); build them with "
                                "DeviceEvaluator.from_model(model, dtype, fold='cublas')")
            if (ev.dtype, ev.plane_stride, getattr(ev, "plane_offset", 0)) != (self.dtype, self.plane_stride, self.plane_offset):
                raise ValueError("all evaluators of a MultiModelEvaluator must share dtype and plane layout")

    def __call__(self, planes: torch.Tensor, row_model: torch.Tensor):
        """row_model: int64 view of the u64 model id per row."""
        pol = a = b = None
        for mid, ev in self.evaluators.items():
            p, x, y = ev(planes)
            key = mid if mid < (1 << 63) else mid - (1 << 64)
            m = row_model == key
            if pol is None:
                pol, a, b = p.float().clone(), x.float().clone(), y.float().clone()
            else:
                pol = torch.where(m[:, None], p.float(), pol)
                a = torch.where(m, x.float(), a)
                b = torch.where(m, y.float(), b)
        return pol, a, b


@dataclass
class RunInfo:
    ticks: int = 0
    wall_s: float = 0.0
    device_s: float = 0.0  # CUDA-event time of the search (first network launch .. last game finished)
    stats: dict = field(default_factory=dict)
    kernel_ms: dict = field(default_factory=dict)  # sampled per-tick device times: k_step, network graph (and k_tail in the python loop)
    engine_bytes: int = 0
    report: dict = field(default_factory=dict)
    n_lanes: int = 1


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap engine-owned device memory."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _wrap_u32(ptr: int, n: int, device) -> torch.Tensor:
    """The engine's uint32 array at `ptr` as an int32 tensor (no copy; counts are < 2^31)."""
    return torch.as_tensor(_DevArray(ptr, n, "<i4"), device=device)


def _wrap_i64(ptr: int, n: int, device) -> torch.Tensor:
    """The engine's uint64 array at `ptr` as an int64 tensor (no copy; same bit patterns)."""
    return torch.as_tensor(_DevArray(ptr, n, "<i8"), device=device)


def _sum_stats(stats: List[dict]) -> dict:
    out = {}
    for s in stats:
        for k, v in s.items():
            out[k] = out.get(k, 0) + v
    return out


class _Lane:
    """One engine + its NN I/O tensors + its stream."""

    def __init__(self, n_slots, max_requests, n_iter, c_expl, c_pen, plane_dtype, device, max_inline, stride, flags,
                 arena_blocks, offset=0, eval_cache_entries=0, spec_rows=0, dirichlet=None):
        self.n_slots = n_slots
        self.offset = offset
        self.engine = Engine(
            n_slots, max(1, max_requests), n_iter, c_expl, c_pen,
            L.PLANES_BF16 if plane_dtype == torch.bfloat16 else L.PLANES_F32, max_inline, device.index, stride, flags,
            arena_blocks, eval_cache_entries, spec_rows, *(dirichlet or (0.0, 0.0)),
        )
        self.io_rows = R = self.engine.io_rows  # n_slots, plus the speculative rows if that is on
        self.planes = torch.zeros(R, stride, dtype=plane_dtype, device=device)
        self.logits = torch.zeros(R, 7, dtype=torch.float32, device=device)
        self.qp = torch.zeros(R, dtype=torch.float32, device=device)
        self.qn = torch.zeros(R, dtype=torch.float32, device=device)
        self.engine.bind_io(self.planes.data_ptr() + offset * self.planes.element_size(), self.logits.data_ptr(),
                            self.qp.data_ptr(), self.qn.data_ptr())
        self.stream = torch.cuda.Stream(device=device)
        self.row_model = _wrap_i64(self.engine.rows_dev()[1], R, device)  # engine-owned, no copy
        self.graphs = {}  # rows -> torch.cuda.CUDAGraph
        self.graph_key = None
        self.pool = None
        self.net = None  # native_net.NativeNet when the evaluator is the library's own kernel

    def attach_native(self, evaluator) -> None:
        """Give this lane a c4a0_net over `evaluator`'s weights: the engine writes its planes straight into
        the net's input buffer, the net writes the engine's logits / q buffers and reads the batch's row
        count on the device."""
        if self.net is not None and self.net.ev is evaluator:
            return
        if self.net is not None:
            self.net.close()
        self.net = evaluator.instantiate(self.io_rows)
        self.net.bind_outputs(self.logits, self.qp, self.qn)
        self.net.bind_row_count(*self.engine.rows_count_dev())
        self.engine.bind_io(self.net.planes_ptr(), self.logits.data_ptr(), self.qp.data_ptr(), self.qn.data_ptr())
        self.planes = self.net.buffer(0)
        self.graphs, self.graph_key = {}, None

    def evaluate(self, evaluator, rows: int) -> None:
        with torch.no_grad():
            if self.net is not None and self.net.ev is evaluator:
                self.net.forward(rows, torch.cuda.current_stream(self.logits.device).cuda_stream)  # rows: read on the device
                return
            if isinstance(evaluator, BuiltinEvaluator):
                # covers every live row whatever `rows` is (the kernel reads the tick's row count)
                self.engine.eval_builtin(evaluator.kind, torch.cuda.current_stream(self.planes.device).cuda_stream)
                return
            if isinstance(evaluator, MultiModelEvaluator):
                pol, a, b = evaluator(self.planes[:rows], self.row_model[:rows])
            elif isinstance(evaluator, DeviceEvaluator):
                evaluator.into(self.planes[:rows], self.logits[:rows], self.qp[:rows], self.qn[:rows])
                return
            else:
                pol, a, b = evaluator(self.planes[:rows])
            self.logits[:rows].copy_(pol.reshape(rows, 7))
            self.qp[:rows].copy_(a.reshape(rows))
            self.qn[:rows].copy_(b.reshape(rows))

    def capture(self, evaluator) -> List[Tuple[int, int]]:
        if self.graph_key is not evaluator:  # identity, and a strong reference: ids can be recycled
            self.graphs, self.graph_key = {}, evaluator
            self.pool = torch.cuda.graph_pool_handle()
            sizes = [b for b in BUCKETS if b < self.io_rows] + [self.io_rows]
            if isinstance(evaluator, BuiltinEvaluator) or (self.net is not None and self.net.ev is evaluator):
                sizes = [self.io_rows]  # one kernel whatever the batch: one graph
            with torch.cuda.stream(self.stream):
                for rows in reversed(sizes):  # largest first: the shared pool is sized once
                    self.evaluate(evaluator, rows)  # eager warm-up (cuBLAS handles, heuristics)
                    self.stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self.pool, stream=self.stream):
                        self.evaluate(evaluator, rows)
                    self.graphs[rows] = g
        return [(rows, self.graphs[rows].raw_cuda_graph_exec()) for rows in sorted(self.graphs)]

    def close(self):
        self.graphs = {}
        self.engine.close()
        if self.net is not None:
            self.net.close()
            self.net = None


class SelfPlaySession:
    """Owns the engine(s) + NN I/O tensors on one GPU; reusable across play() calls."""

    def __init__(
        self,
        n_slots: int,
        max_requests: int,
        n_mcts_iterations: int,
        c_exploration: float,
        c_ply_penalty: float,
        plane_dtype: torch.dtype = torch.float32,
        device: int = 0,
        max_inline_sims: int = 0,
        plane_stride: int = 84,
        plane_offset: int = 0,
        n_lanes: Optional[int] = None,
        dedup: Optional[bool] = None,
        arena_blocks: Optional[int] = None,
        eval_cache: bool = False,
        eval_cache_entries: int = 0,
        speculate: bool = False,
        spec_rows: int = 0,
        dirichlet: Optional[Tuple[float, float]] = None,
    ):
        """`eval_cache=True` lets the engine answer repeated (position, model) leaves of one play() call from
        a device table instead of the network; only for evaluators that are pure functions of the position."""
        if not torch.cuda.is_available():
            raise RuntimeError("c4a0_b200 needs a CUDA device: there is no CPU fallback")
        if plane_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("plane_dtype must be float32 or bfloat16")
        n_lanes = DEFAULTS["n_lanes"] if n_lanes is None else n_lanes
        max_inline_sims = max_inline_sims or DEFAULTS["max_inline_sims"]
        dedup = DEFAULTS["dedup"] if dedup is None else dedup
        if n_slots < 2 * 256:
            n_lanes = 1  # tiny batches: nothing to overlap
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.n_slots = n_slots
        self.plane_dtype = plane_dtype
        self.plane_stride = plane_stride
        self.plane_offset = plane_offset
        if speculate and not eval_cache:
            raise ValueError("speculate=True needs eval_cache=True")
        flags = ((0 if dedup else L.FLAG_NO_DEDUP) | (L.FLAG_EVAL_CACHE if eval_cache else 0)
                 | (L.FLAG_SPECULATE if speculate else 0))
        self.eval_cache = bool(eval_cache)
        self.speculate = bool(speculate)
        if arena_blocks is None:
            arena_blocks = DEFAULTS["arena_blocks"]
        if arena_blocks is None:
            # Roomy arena halves make re-rooting copy-free (see engine.cu).  Default 8x the minimum: 34 GB
            # for 16,384 games x 600 sims, which leaves the GPU to a trainer as well; a half then fills up
            # a few thousand times per job and its compaction costs ~3 % of a bench step (32x = 110 GB has
            # none: 1,410 vs 1,455 ms, profiles/r01_summary.md).  Never more than 35 % of the free memory.
            # 2 halves x 160 B per block per game.
            free_b, _ = torch.cuda.mem_get_info(self.device)
            minimum = n_mcts_iterations + 2
            arena_blocks = max(minimum, min(8 * minimum, int(free_b * 0.35) // (n_slots * 320)))
        per = [(n_slots + i) // n_lanes for i in range(n_lanes)][::-1]
        self.lanes = [
            _Lane(s, max_requests, n_mcts_iterations, c_exploration, c_ply_penalty, plane_dtype, self.device,
                  max_inline_sims, plane_stride, flags, arena_blocks, plane_offset, eval_cache_entries, spec_rows,
                  dirichlet if dirichlet is not None else DEFAULTS["dirichlet"])
            for s in per if s > 0
        ]

    # single-lane conveniences used by tests
    @property
    def engine(self) -> Engine:
        return self.lanes[0].engine

    @property
    def planes(self):
        return self.lanes[0].planes

    @property
    def logits(self):
        return self.lanes[0].logits

    @property
    def qp(self):
        return self.lanes[0].qp

    @property
    def qn(self):
        return self.lanes[0].qn

    @property
    def stream(self):
        return self.lanes[0].stream

    def close(self):
        for ln in self.lanes:
            ln.close()

    def _split(self, n_req: int) -> List[Tuple[int, int]]:
        """Contiguous request ranges per lane, proportional to the lanes' slot counts."""
        total = sum(ln.n_slots for ln in self.lanes)
        out, lo = [], 0
        for i, ln in enumerate(self.lanes):
            hi = n_req if i == len(self.lanes) - 1 else lo + (n_req * ln.n_slots) // total
            out.append((lo, hi))
            lo = hi
        return out

    # ------------------------------------------------------------------------------------------
    def play(
        self,
        game_id: Sequence[int],
        player0_id: Sequence[int],
        player1_id: Sequence[int],
        evaluator: Callable,
        host_loop: Optional[str] = None,
        poll_every: Optional[int] = None,
        sample_kernels_every: Optional[int] = None,
        fetch: bool = True,
    ) -> Tuple[Optional[GameSamples], RunInfo]:
        """Play all requested games; returns host samples (request order) and run information.

        host_loop = "native": bucketed CUDA graphs driven by the C++ loop (c4a0_engine_run);
                    "python": eager network calls and engine.step() from Python (any callable)."""
        want = (getattr(evaluator, "plane_stride", self.plane_stride), getattr(evaluator, "plane_offset", self.plane_offset))
        if want != (self.plane_stride, self.plane_offset):
            raise ValueError(f"evaluator expects plane layout (stride, offset) = {want}, the session was built with "
                             f"{(self.plane_stride, self.plane_offset)}")
        host_loop = DEFAULTS["host_loop"] if host_loop is None else host_loop
        poll_every = DEFAULTS["poll_every"] if poll_every is None else poll_every
        sample_kernels_every = DEFAULTS["sample_kernels_every"] if sample_kernels_every is None else sample_kernels_every
        gid = np.ascontiguousarray(game_id, dtype=np.uint64)
        p0 = np.ascontiguousarray(player0_id, dtype=np.uint64)
        p1 = np.ascontiguousarray(player1_id, dtype=np.uint64)
        n_req = len(gid)
        info = RunInfo(engine_bytes=sum(ln.engine.device_bytes for ln in self.lanes), n_lanes=len(self.lanes))
        t0 = time.perf_counter()
        ranges = self._split(n_req)
        self._last_ranges = ranges
        from .native_net import NativeEvaluator

        if isinstance(evaluator, NativeEvaluator):
            for ln in self.lanes:
                ln.attach_native(evaluator)  # before set_requests(): it packs the planes of the initial roots
        nvtx = torch.cuda.nvtx  # ranges for nsys / ncu --nvtx timelines; no-ops without a profiler
        nvtx.range_push("c4a0.set_requests")
        for ln, (lo, hi) in zip(self.lanes, ranges):
            ln.engine.set_requests(gid[lo:hi], p0[lo:hi], p1[lo:hi], ln.stream.cuda_stream)
        nvtx.range_pop()
        if host_loop == "native":
            native = isinstance(evaluator, NativeEvaluator) and all(ln.net is not None and ln.net.ev is evaluator for ln in self.lanes)
            if not native:
                nvtx.range_push("c4a0.capture_graphs")
                graphs = [ln.capture(evaluator) for ln in self.lanes]
                nvtx.range_pop()
            nvtx.range_push("c4a0.search")
            try:
                if native:  # the library's network kernel: launched by the C++ loop itself, no graphs
                    rep = run_engines_net([ln.engine for ln in self.lanes], [ln.net._h.value for ln in self.lanes],
                                          [ln.stream.cuda_stream for ln in self.lanes], 0, sample_kernels_every)
                else:
                    rep = run_engines(
                        [ln.engine for ln in self.lanes], graphs, [ln.stream.cuda_stream for ln in self.lanes], 0,
                        sample_kernels_every,
                    )
            finally:
                nvtx.range_pop()
            info.report = rep
            info.ticks = int(rep["ticks"])
            info.device_s = rep["device_ms"] / 1e3
            if rep["kernel_samples"]:
                n = rep["kernel_samples"]
                info.kernel_ms = {"k_step": rep["k_step_ms_sum"] / n, "nn": rep["nn_ms_sum"] / n, "samples": n}
        elif host_loop == "python":
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(self.lanes[0].stream)
            ks, km, kn, ticks = 0.0, 0.0, 0, 0
            live = [hi > lo for lo, hi in ranges]
            sizes = {id(ln): [b for b in BUCKETS if b < ln.io_rows] + [ln.io_rows] for ln in self.lanes}
            rows = {id(ln): ln.engine.poll(ln.stream.cuda_stream).n_rows for ln in self.lanes}
            while any(live):
                for i, ln in enumerate(self.lanes):
                    if not live[i]:
                        continue
                    with torch.cuda.stream(ln.stream):
                        # the smallest bucket covering the live rows, like the native loop
                        ln.evaluate(evaluator, next(b for b in sizes[id(ln)] if b >= rows[id(ln)]))
                        if sample_kernels_every and ticks % sample_kernels_every == 0:
                            x, y = ln.engine.step_timed(ln.stream.cuda_stream)
                            ks, km, kn = ks + x, km + y, kn + 1
                        else:
                            ln.engine.step(ln.stream.cuda_stream)
                    ticks += 1
                    p = ln.engine.poll(ln.stream.cuda_stream)
                    rows[id(ln)] = p.n_rows
                    live[i] = p.n_finished < p.n_requests
            for ln in self.lanes:
                ln.stream.synchronize()
            ev1.record(self.lanes[0].stream)
            ev1.synchronize()
            info.device_s = ev0.elapsed_time(ev1) / 1e3
            info.ticks = ticks
            if kn:
                info.kernel_ms = {"k_step": ks / kn, "k_tail": km / kn, "samples": kn}
        else:
            raise ValueError("host_loop must be 'native' or 'python'")
        info.stats = _sum_stats([ln.engine.stats(ln.stream.cuda_stream) for ln in self.lanes])
        out = None
        if fetch:
            nvtx.range_push("c4a0.fetch_results")
            parts = [ln.engine.fetch_results(0, hi - lo, ln.stream.cuda_stream) for ln, (lo, hi) in zip(self.lanes, ranges)]
            nvtx.range_pop()
            out = GameSamples(*[
                np.concatenate([getattr(p, f) for p in parts])
                for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty")
            ])
        info.wall_s = time.perf_counter() - t0
        return out, info

    def export_tensors(self, augment: bool = True):
        """All samples of the last play() as CUDA tensors (pos [S,2,6,7], policy [S,7], q_penalty [S],
        q_no_penalty [S]) without leaving the device; augment=True appends the mirror images
        (training.py:317-333 does both on the host, one Python object per sample)."""
        outs = []
        for ln, (lo, hi) in zip(self.lanes, self._last_ranges):
            n = hi - lo
            if n == 0:
                continue
            with torch.cuda.stream(ln.stream):
                ptrs = ln.engine.results_dev()
                # the sample counts live in the engine's store: wrap them without a copy
                counts = _wrap_u32(ptrs[0], n, self.device).to(torch.int64)
                offs = (torch.cumsum(counts, 0) - counts).to(torch.int32).contiguous()
                total = int(counts.sum().item())
                rows = total * (2 if augment else 1)
                pos = torch.empty(rows, 2, 6, 7, dtype=torch.float32, device=self.device)
                pol = torch.empty(rows, 7, dtype=torch.float32, device=self.device)
                qp = torch.empty(rows, dtype=torch.float32, device=self.device)
                qn = torch.empty(rows, dtype=torch.float32, device=self.device)
                ln.engine.export_samples(0, n, offs.data_ptr(), total, augment, pos.data_ptr(), pol.data_ptr(),
                                         qp.data_ptr(), qn.data_ptr(), ln.stream.cuda_stream)
                ln.stream.synchronize()
            outs.append((pos, pol, qp, qn))
        if not outs:
            z = torch.zeros(0, device=self.device)
            return z.reshape(0, 2, 6, 7), z.reshape(0, 7), z, z
        return tuple(torch.cat([o[i] for o in outs]) for i in range(4))

    # ------------------------------------------------------------------------------------------
    def play_callback(
        self,
        game_id: Sequence[int],
        player0_id: Sequence[int],
        player1_id: Sequence[int],
        cb: Callable,
        max_nn_batch_size: int,
    ) -> Tuple[GameSamples, RunInfo]:
        """The reference's numpy-callback contract (pybridge.rs:161-199).  The engine has already
        de-duplicated the waiting leaves per (position, model) (self_play.rs:203-208); every tick
        the live rows are grouped by model id, cut into batches of at most max_nn_batch_size and
        handed to `cb` on the host."""
        if self.plane_dtype != torch.float32:
            raise ValueError("the numpy callback path needs float32 planes")
        if len(self.lanes) != 1:
            raise ValueError("the numpy callback path runs on one lane (n_lanes=1)")
        ln = self.lanes[0]
        info = RunInfo(engine_bytes=ln.engine.device_bytes)
        n_req = len(game_id)
        self._last_ranges = [(0, n_req)]
        t0 = time.perf_counter()
        S = ln.io_rows
        h_logits = torch.zeros(S, 7, dtype=torch.float32).pin_memory()
        h_qp = torch.zeros(S, dtype=torch.float32).pin_memory()
        h_qn = torch.zeros(S, dtype=torch.float32).pin_memory()
        with torch.cuda.stream(ln.stream):
            s = ln.stream.cuda_stream
            ln.engine.set_requests(game_id, player0_id, player1_id, s)
            ticks = 0
            while n_req:
                n_rows, _, _, model = ln.engine.fetch_rows(s)
                if n_rows:
                    o = ln.offset
                    planes = ln.planes[:n_rows, o : o + 84].cpu().numpy().reshape(n_rows, 2, 6, 7)
                    lg, a, b = h_logits.numpy(), h_qp.numpy(), h_qn.numpy()
                    for mid in np.unique(model):
                        rows = np.nonzero(model == mid)[0]
                        for lo in range(0, rows.size, max_nn_batch_size):
                            sel = rows[lo : lo + max_nn_batch_size]
                            batch = np.ascontiguousarray(planes[sel])
                            pol, qa, qb = cb(int(mid), batch)
                            pol = np.asarray(pol, dtype=np.float32)
                            if pol.shape != (sel.size, 7):
                                raise ValueError(f"callback returned policy of shape {pol.shape}, expected {(sel.size, 7)}")
                            lg[sel] = pol
                            a[sel] = np.asarray(qa, dtype=np.float32).reshape(sel.size)
                            b[sel] = np.asarray(qb, dtype=np.float32).reshape(sel.size)
                    ln.logits[:n_rows].copy_(h_logits[:n_rows], non_blocking=True)
                    ln.qp[:n_rows].copy_(h_qp[:n_rows], non_blocking=True)
                    ln.qn[:n_rows].copy_(h_qn[:n_rows], non_blocking=True)
                ln.engine.step(s)
                ticks += 1
                if ln.engine.poll(s).n_finished >= n_req:
                    break
            info.ticks = ticks
            info.stats = ln.engine.stats(s)
            out = ln.engine.fetch_results(0, n_req, s)
        info.wall_s = time.perf_counter() - t0
        return out, info
