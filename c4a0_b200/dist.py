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
    """Broadcast parameters and buffers (BatchNorm statistics) from `src` as ONE flat tensor per dtype
    (a c4a0 network is ~30 tensors; one NCCL call instead of thirty); returns bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    n = 0
    with torch.no_grad():
        groups = {}
        for t in list(model.parameters()) + list(model.buffers()):
            groups.setdefault((t.dtype, t.device), []).append(t.data)
        for tensors in groups.values():
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.broadcast(flat, src=src)
            o = 0
            for t in tensors:
                t.copy_(flat[o : o + t.numel()].view_as(t))
                o += t.numel()
            n += flat.numel() * flat.element_size()
    return n


# ------------------------------------------------------------------------------------------------
# Sample gather: only the valid samples travel, packed, rank -> dst point to point
# ------------------------------------------------------------------------------------------------
PACK_WORDS = 13  # int32 words per sample: mask (2), value (2), policy (7), q_penalty, q_no_penalty


def pack_samples(n_samples: torch.Tensor, mask: torch.Tensor, value: torch.Tensor, policy: torch.Tensor,
                 q_penalty: torch.Tensor, q_no_penalty: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G] counts + padded [G,43,...] sample arrays (int64 bit patterns for the u64 fields) -> (counts int32 [G],
    packed int32 [S, 13]) holding the S valid samples in game order.  Runs where the tensors live: on the
    engine's device sample store (no host copy) or on CPU tensors (gloo tests)."""
    G = n_samples.shape[0]
    counts = n_samples.to(torch.int32)
    valid = torch.arange(43, device=n_samples.device)[None, :] < counts[:, None]
    words = torch.cat([
        mask.reshape(G, 43, 1).view(torch.int32).reshape(G, 43, 2),
        value.reshape(G, 43, 1).view(torch.int32).reshape(G, 43, 2),
        policy.reshape(G, 43, 7).view(torch.int32),
        q_penalty.reshape(G, 43, 1).view(torch.int32),
        q_no_penalty.reshape(G, 43, 1).view(torch.int32),
    ], dim=2)
    return counts, words[valid].contiguous()


def unpack_samples(counts: np.ndarray, packed: np.ndarray) -> GameSamples:
    """Inverse of pack_samples on the host: padded [G,43,...] arrays (unused cells zero)."""
    counts = np.asarray(counts, dtype=np.int64)
    G = len(counts)
    words = np.zeros((G, 43, PACK_WORDS), np.int32)
    valid = np.arange(43)[None, :] < counts[:, None]
    words[valid] = np.asarray(packed, dtype=np.int32).reshape(-1, PACK_WORDS)
    return GameSamples(
        counts.astype(np.uint32),
        np.ascontiguousarray(words[:, :, 0:2]).view(np.uint64).reshape(G, 43),
        np.ascontiguousarray(words[:, :, 2:4]).view(np.uint64).reshape(G, 43),
        np.ascontiguousarray(words[:, :, 4:11]).view(np.float32),
        np.ascontiguousarray(words[:, :, 11]).view(np.float32),
        np.ascontiguousarray(words[:, :, 12]).view(np.float32),
    )


def gather_packed(meta: torch.Tensor, counts: torch.Tensor, packed: torch.Tensor, dst: int = 0):
    """Send (meta int64 [G,3], counts int32 [G], packed int32 [S,13]) of every rank to `dst`: one small
    all_gather of the sizes, then point-to-point transfers of exactly the valid bytes (NCCL send/recv over
    NVLink, or gloo).  On `dst`: lists of the three tensors in rank order; elsewhere None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = packed.device
    sizes = torch.tensor([meta.shape[0], packed.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [tuple(int(v) for v in t.tolist()) for t in all_sizes]
    if rank != dst:
        ops = [dist.P2POp(dist.isend, t.contiguous(), dst) for t in (meta, counts, packed) if t.numel()]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None
    out, ops = [], []
    for r, (g, s_) in enumerate(all_sizes):
        if r == dst:
            out.append((meta, counts, packed))
            continue
        bufs = (torch.empty(g, 3, dtype=torch.int64, device=dev), torch.empty(g, dtype=torch.int32, device=dev),
                torch.empty(s_, PACK_WORDS, dtype=torch.int32, device=dev))
        out.append(bufs)
        ops += [dist.P2POp(dist.irecv, t, r) for t in bufs if t.numel()]
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    return out


_PINNED = {}  # (name, dtype) -> flat pinned staging tensor, grown on demand (cudaHostAlloc of ~100 MB costs tens of ms per call)


def _to_host(t: torch.Tensor, name: str = "", copy: bool = True) -> np.ndarray:
    """Device tensor -> numpy array, staged through a cached pinned buffer (a pageable destination halves the
    copy rate; allocating the pinned buffer anew every generation costs more than the copy).  copy=False returns
    a view of the staging buffer `name`: valid until the next call with the same name."""
    if not t.is_cuda:
        return t.numpy()
    n = t.numel()
    key = (name, t.dtype)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < n:
        buf = _PINNED[key] = torch.empty(max(n, 1024), dtype=t.dtype, pin_memory=True)
    h = buf[:n].view(t.shape)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy().copy() if copy else h.numpy()


def unpack_samples_device(counts: torch.Tensor, packed: torch.Tensor):
    """pack_samples' inverse where the tensors live (rank 0's GPU): padded [G,43,...] tensors."""
    G = counts.shape[0]
    words = torch.zeros(G, 43, PACK_WORDS, dtype=torch.int32, device=packed.device)
    valid = torch.arange(43, device=counts.device)[None, :] < counts[:, None]
    words[valid] = packed
    return (counts, words[:, :, 0:2].contiguous().view(torch.int64).reshape(G, 43), words[:, :, 2:4].contiguous().view(torch.int64).reshape(G, 43),
            words[:, :, 4:11].contiguous().view(torch.float32), words[:, :, 11].contiguous().view(torch.float32),
            words[:, :, 12].contiguous().view(torch.float32))


def _finish_gather(parts, packed_result: bool = False):
    """Rank-ordered parts on dst -> host.  packed_result: (meta u64 [G,3], counts i32 [G], packed i32 [S,13]) —
    exactly the valid samples, 52 bytes each; otherwise (meta, GameSamples) with padded [G,43,...] arrays,
    unpacked on the device when there is one."""
    meta = torch.cat([p[0] for p in parts])
    counts = torch.cat([p[1] for p in parts])
    packed = torch.cat([p[2] for p in parts])
    if packed_result:  # views of the pinned staging buffers: valid until the next gather (see gather_session_samples)
        return (_to_host(meta, "meta", copy=False).view(np.uint64), _to_host(counts, "counts", copy=False),
                _to_host(packed, "packed", copy=False))
    if packed.is_cuda:
        c, m, v, pol, qp, qn = unpack_samples_device(counts, packed)
        return _to_host(meta).view(np.uint64), GameSamples(_to_host(c).astype(np.uint32), _to_host(m).view(np.uint64), _to_host(v).view(np.uint64),
                                                           _to_host(pol), _to_host(qp), _to_host(qn))
    return meta.numpy().view(np.uint64), unpack_samples(counts.numpy(), packed.numpy())


def gather_samples(meta: np.ndarray, soa: GameSamples, device: Optional[torch.device] = None, dst: int = 0):
    """Concatenate every rank's finished games on `dst` in rank order, from HOST arrays.  Returns (meta, soa)
    on `dst` and (None, None) elsewhere.  (gather_session_samples() does the same from the engines' device
    sample stores without the host round trip.)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return meta, soa
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def up(a: np.ndarray, view=None) -> torch.Tensor:
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a.view(view) if view is not None else a).to(device)

    counts, packed = pack_samples(up(soa.n_samples.astype(np.int32)), up(soa.mask, np.int64), up(soa.value, np.int64),
                                  up(soa.policy), up(soa.q_penalty), up(soa.q_no_penalty))
    parts = gather_packed(up(np.ascontiguousarray(meta, dtype=np.uint64).reshape(-1, 3), np.int64), counts, packed, dst)
    if parts is None:
        return None, None
    return _finish_gather(parts)


def gather_session_samples(session, meta: np.ndarray, dst: int = 0, packed_result: bool = False):
    """The samples of `session`'s last play() from every rank onto `dst`, packed on the device straight out of
    the engines' sample stores (c4a0_engine_results_dev): no padded arrays, no host bounce on the senders.
    Returns (meta u64 [G,3], GameSamples) on `dst` — or, with packed_result, (meta, counts, packed [S,13]) as views
    of pinned staging buffers that the next packed gather overwrites (copy what must outlive it) — and a tuple of
    Nones elsewhere."""
    from .selfplay import _wrap_i64, _wrap_u32  # engine-owned device arrays as tensors, no copy

    dev = session.device
    cs, ps = [], []
    for ln, (lo, hi) in zip(session.lanes, session._last_ranges):
        n = hi - lo
        if n == 0:
            continue
        with torch.cuda.stream(ln.stream):
            p_n, p_mask, p_value, p_pol, p_qp, p_qn = ln.engine.results_dev()
            c, p = pack_samples(
                _wrap_u32(p_n, n, dev), _wrap_i64(p_mask, n * 43, dev).view(n, 43), _wrap_i64(p_value, n * 43, dev).view(n, 43),
                _wrap_u32(p_pol, n * 43 * 7, dev).view(torch.float32).view(n, 43, 7),
                _wrap_u32(p_qp, n * 43, dev).view(torch.float32).view(n, 43),
                _wrap_u32(p_qn, n * 43, dev).view(torch.float32).view(n, 43))
            ln.stream.synchronize()
        cs.append(c)
        ps.append(p)
    counts = torch.cat(cs) if cs else torch.zeros(0, dtype=torch.int32, device=dev)
    packed = torch.cat(ps) if ps else torch.zeros(0, PACK_WORDS, dtype=torch.int32, device=dev)
    m = torch.from_numpy(np.ascontiguousarray(meta, dtype=np.uint64).reshape(-1, 3).view(np.int64)).to(dev)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return _finish_gather([(m, counts, packed)], packed_result)
    parts = gather_packed(m, counts, packed, dst)
    if parts is None:
        return (None, None, None) if packed_result else (None, None)
    return _finish_gather(parts, packed_result)
