"""Data-parallel game sharding: one process per GPU, games split by contiguous request ranges.

The reference has no multi-GPU path at all — self-play games are independent
(each `MctsGame` owns its tree and its RNG is seeded by `game_id`, rust/src/mcts.rs:215), so the
only communication a generation needs is
  * a broadcast of the new weights from rank 0 (NCCL over NVLink), and
  * a gather of the finished samples to rank 0,
with no collective on the search path.  Results per game_id are identical to a 1-GPU run.
Works with the `gloo` backend on CPU tensors too (used by the CPU test-suite).
"""

from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import GameSamples


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* (torchrun). Returns
    (rank, world_size, local_rank); a no-op single process when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous request range [lo, hi) of `rank`: sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_model(model: torch.nn.Module, src: int = 0) -> int:
    """Broadcast parameters and buffers (BatchNorm statistics) from `src`; returns bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
            n += t.numel() * t.element_size()
    return n


def _gather_rows(x: torch.Tensor, counts: List[int], dst: int) -> Optional[torch.Tensor]:
    """Gather variable-length leading-dim tensors to `dst` (padded all_gather, then trimmed)."""
    world = dist.get_world_size()
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if dist.get_rank() != dst:
        return None
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def gather_samples(meta: np.ndarray, soa: GameSamples, device: Optional[torch.device] = None, dst: int = 0):
    """Concatenate every rank's finished games on `dst` in rank order.  Returns (meta, soa) on
    `dst` and (None, None) elsewhere.  uint64 fields travel as int64 bit patterns."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return meta, soa
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    world = dist.get_world_size()
    n_local = torch.tensor([len(soa.n_samples)], dtype=torch.int64, device=device)
    all_n = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(all_n, n_local)
    counts = [int(t.item()) for t in all_n]

    def up(a: np.ndarray, view=None) -> torch.Tensor:
        a = np.ascontiguousarray(a)
        if view is not None:
            a = a.view(view)
        return torch.from_numpy(a).to(device)

    parts = dict(
        meta=up(np.ascontiguousarray(meta, dtype=np.uint64).reshape(-1, 3), np.int64),
        n_samples=up(soa.n_samples.astype(np.int64)),
        mask=up(soa.mask, np.int64),
        value=up(soa.value, np.int64),
        policy=up(soa.policy),
        q_penalty=up(soa.q_penalty),
        q_no_penalty=up(soa.q_no_penalty),
    )
    got = {k: _gather_rows(v, counts, dst) for k, v in parts.items()}
    if dist.get_rank() != dst:
        return None, None
    h = {k: v.cpu().numpy() for k, v in got.items()}
    out = GameSamples(
        h["n_samples"].astype(np.uint32),
        h["mask"].view(np.uint64),
        h["value"].view(np.uint64),
        h["policy"],
        h["q_penalty"],
        h["q_no_penalty"],
    )
    return h["meta"].view(np.uint64), out
