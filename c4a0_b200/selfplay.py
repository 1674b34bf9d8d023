"""Lockstep self-play driver: the engine's tree kernels + a PyTorch network on one CUDA stream.

This is the B200 replacement for the thread/channel machinery of rust/src/self_play.rs:39-246
(`self_play`, `NNThread`).  There the NN thread collects leaf positions from a queue, builds a
numpy batch, calls Python, and fans results back through another queue; here a *tick* is

    network(planes) -> logits, q_penalty, q_no_penalty        (PyTorch, bf16 or fp32)
    engine.step()   -> expand + backup + move + select, writes the next planes   (our kernels)

on the same device buffers, captured once into a CUDA graph and replayed until every game has
finished.  The host only polls a progress counter every `poll_every` ticks.

Two evaluator contracts are supported:
  * device evaluators (fast path): `fn(planes: cuda Tensor[B,2,6,7]) -> (policy[B,7], qp[B], qn[B])`
    cuda tensors — e.g. a `c4a0_b200.nn.ConnectFourNet` wrapped in `DeviceEvaluator`;
  * the reference's numpy callback `cb(model_id, ndarray[B,2,6,7]) -> (policy, qp, qn)`
    (rust/src/pybridge.rs:161-199), served by `run_callback` with host copies every tick.
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import Engine, GameSamples


# process-wide defaults, overridable by tools (bench.py turns kernel sampling on)
DEFAULTS = {"use_cuda_graph": True, "poll_every": 64, "sample_kernels_every": 0}


class DeviceEvaluator:
    """Marks a callable as device-capable: planes (cuda tensor) in, three cuda tensors out."""

    def __init__(self, module_or_fn, dtype: torch.dtype = torch.float32):
        self.fn = module_or_fn
        self.dtype = dtype
        if isinstance(module_or_fn, torch.nn.Module):
            module_or_fn.eval()

    def __call__(self, planes: torch.Tensor):
        return self.fn(planes)


@dataclass
class RunInfo:
    ticks: int = 0
    wall_s: float = 0.0
    device_s: float = 0.0  # CUDA-event time of the tick loop (first tick .. all games finished)
    stats: dict = field(default_factory=dict)
    kernel_ms: dict = field(default_factory=dict)  # sampled k_step / k_move durations
    engine_bytes: int = 0


class SelfPlaySession:
    """Owns one engine + its NN I/O tensors on one GPU; reusable across play() calls."""

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
    ):
        if not torch.cuda.is_available():
            raise RuntimeError("c4a0_b200 needs a CUDA device: there is no CPU fallback")
        if plane_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("plane_dtype must be float32 or bfloat16")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.n_slots = n_slots
        self.plane_dtype = plane_dtype
        self.engine = Engine(
            n_slots, max_requests, n_mcts_iterations, c_exploration, c_ply_penalty,
            L.PLANES_BF16 if plane_dtype == torch.bfloat16 else L.PLANES_F32, max_inline_sims, device,
        )
        self.planes = torch.zeros(n_slots, 2, 6, 7, dtype=plane_dtype, device=self.device)
        self.logits = torch.zeros(n_slots, 7, dtype=torch.float32, device=self.device)
        self.qp = torch.zeros(n_slots, dtype=torch.float32, device=self.device)
        self.qn = torch.zeros(n_slots, dtype=torch.float32, device=self.device)
        self.engine.bind_io(self.planes.data_ptr(), self.logits.data_ptr(), self.qp.data_ptr(), self.qn.data_ptr())
        self.stream = torch.cuda.Stream(device=self.device)
        self._graph = None
        self._graph_key = None

    def close(self):
        self._graph = None
        self.engine.close()

    # ------------------------------------------------------------------------------------------
    def _tick(self, evaluator) -> None:
        with torch.no_grad():
            pol, a, b = evaluator(self.planes)
            self.logits.copy_(pol.reshape(self.n_slots, 7))
            self.qp.copy_(a.reshape(self.n_slots))
            self.qn.copy_(b.reshape(self.n_slots))
        self.engine.step(torch.cuda.current_stream().cuda_stream)

    def play(
        self,
        game_id: Sequence[int],
        player0_id: Sequence[int],
        player1_id: Sequence[int],
        evaluator: Callable,
        use_cuda_graph: Optional[bool] = None,
        poll_every: Optional[int] = None,
        sample_kernels_every: Optional[int] = None,
        fetch: bool = True,
    ) -> Tuple[Optional[GameSamples], RunInfo]:
        """Play all requested games; returns host samples (request order) and run information."""
        use_cuda_graph = DEFAULTS["use_cuda_graph"] if use_cuda_graph is None else use_cuda_graph
        poll_every = DEFAULTS["poll_every"] if poll_every is None else poll_every
        sample_kernels_every = DEFAULTS["sample_kernels_every"] if sample_kernels_every is None else sample_kernels_every
        info = RunInfo(engine_bytes=self.engine.device_bytes)
        n_req = len(game_id)
        t0 = time.perf_counter()
        with torch.cuda.stream(self.stream):
            s = self.stream.cuda_stream
            self.engine.set_requests(game_id, player0_id, player1_id, s)
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(self.stream)
            ticks = 0
            finished = n_req == 0
            graph = None
            if use_cuda_graph and not finished:
                key = id(evaluator)
                if self._graph is None or self._graph_key != key:
                    for _ in range(3):  # real ticks; also warms cuBLAS/cuDNN before capture
                        self._tick(evaluator)
                        ticks += 1
                    self.stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._tick(evaluator)
                    self._graph, self._graph_key = g, key
                graph = self._graph
            ks, km, kn = 0.0, 0.0, 0
            while not finished:
                for i in range(poll_every):
                    if sample_kernels_every and (ticks % sample_kernels_every) == 0:
                        # an eager tick whose two engine kernels are bracketed by CUDA events
                        with torch.no_grad():
                            pol, a, b = evaluator(self.planes)
                            self.logits.copy_(pol.reshape(self.n_slots, 7))
                            self.qp.copy_(a.reshape(self.n_slots))
                            self.qn.copy_(b.reshape(self.n_slots))
                        x, y = self.engine.step_timed(s)
                        ks, km, kn = ks + x, km + y, kn + 1
                    elif graph is not None:
                        graph.replay()
                    else:
                        self._tick(evaluator)
                    ticks += 1
                p = self.engine.poll(s)
                finished = p.n_finished >= n_req
            ev1.record(self.stream)
            self.stream.synchronize()
            info.device_s = ev0.elapsed_time(ev1) / 1e3
            info.ticks = ticks
            info.stats = self.engine.stats(s)
            if kn:
                info.kernel_ms = {"k_step": ks / kn, "k_move": km / kn, "samples": kn}
            out = self.engine.fetch_results(0, n_req, s) if fetch else None
        info.wall_s = time.perf_counter() - t0
        return out, info

    # ------------------------------------------------------------------------------------------
    def play_callback(
        self,
        game_id: Sequence[int],
        player0_id: Sequence[int],
        player1_id: Sequence[int],
        cb: Callable,
        max_nn_batch_size: int,
    ) -> Tuple[GameSamples, RunInfo]:
        """The reference's numpy-callback contract (pybridge.rs:161-199): every tick the waiting
        leaf positions are grouped by model id, de-duplicated (self_play.rs:203-208), cut into
        batches of at most max_nn_batch_size and handed to `cb` on the host."""
        if self.plane_dtype != torch.float32:
            raise ValueError("the numpy callback path needs float32 planes")
        info = RunInfo(engine_bytes=self.engine.device_bytes)
        n_req = len(game_id)
        t0 = time.perf_counter()
        S = self.n_slots
        h_logits = torch.zeros(S, 7, dtype=torch.float32).pin_memory()
        h_qp = torch.zeros(S, dtype=torch.float32).pin_memory()
        h_qn = torch.zeros(S, dtype=torch.float32).pin_memory()
        with torch.cuda.stream(self.stream):
            s = self.stream.cuda_stream
            self.engine.set_requests(game_id, player0_id, player1_id, s)
            ticks = 0
            while n_req:
                state, mask, value, model = self.engine.fetch_rows(s)
                waiting = np.nonzero(state == L.ROW_WAIT_NN)[0]
                if waiting.size:
                    planes = self.planes.cpu().numpy()
                    lg, a, b = h_logits.numpy(), h_qp.numpy(), h_qn.numpy()
                    for mid in np.unique(model[waiting]):
                        rows = waiting[model[waiting] == mid]
                        keys = np.stack([mask[rows], value[rows]], axis=1)
                        _, first, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
                        inverse = inverse.reshape(-1)
                        u_pol = np.empty((first.size, 7), np.float32)
                        u_a = np.empty(first.size, np.float32)
                        u_b = np.empty(first.size, np.float32)
                        for lo in range(0, first.size, max_nn_batch_size):
                            sel = first[lo : lo + max_nn_batch_size]
                            batch = np.ascontiguousarray(planes[rows[sel]])
                            pol, qa, qb = cb(int(mid), batch)
                            pol = np.asarray(pol, dtype=np.float32)
                            if pol.shape != (sel.size, 7):
                                raise ValueError(f"callback returned policy of shape {pol.shape}, expected {(sel.size, 7)}")
                            u_pol[lo : lo + sel.size] = pol
                            u_a[lo : lo + sel.size] = np.asarray(qa, dtype=np.float32).reshape(sel.size)
                            u_b[lo : lo + sel.size] = np.asarray(qb, dtype=np.float32).reshape(sel.size)
                        lg[rows] = u_pol[inverse]
                        a[rows] = u_a[inverse]
                        b[rows] = u_b[inverse]
                    self.logits.copy_(h_logits, non_blocking=True)
                    self.qp.copy_(h_qp, non_blocking=True)
                    self.qn.copy_(h_qn, non_blocking=True)
                self.engine.step(s)
                ticks += 1
                if self.engine.poll(s).n_finished >= n_req:
                    break
            info.ticks = ticks
            info.stats = self.engine.stats(s)
            out = self.engine.fetch_results(0, n_req, s)
        info.wall_s = time.perf_counter() - t0
        return out, info
