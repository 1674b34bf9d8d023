"""world_size-2 `gloo` test of the data-parallel plumbing (c4a0_b200/dist.py): request sharding,
weight broadcast, sample gather.  No GPU."""

import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_samples(lo, hi):
    from c4a0_b200.engine import GameSamples

    n = hi - lo
    ids = np.arange(lo, hi)
    soa = GameSamples(
        (ids % 5 + 8).astype(np.uint32),
        (ids[:, None] * 1000 + np.arange(43)[None, :]).astype(np.uint64) | np.uint64(1 << 63),
        (ids[:, None] * 7 + np.arange(43)[None, :]).astype(np.uint64),
        np.broadcast_to((ids[:, None, None] + np.arange(7)[None, None, :] / 8).astype(np.float32), (n, 43, 7)).copy(),
        np.broadcast_to((ids[:, None] / 16).astype(np.float32), (n, 43)).copy(),
        np.broadcast_to((-ids[:, None] / 32).astype(np.float32), (n, 43)).copy(),
    )
    # cells past a game's sample count are zero, as the engine leaves them (only valid samples travel)
    dead = np.arange(43)[None, :] >= soa.n_samples[:, None]
    for a in (soa.mask, soa.value, soa.policy, soa.q_penalty, soa.q_no_penalty):
        a[dead] = 0
    meta = np.stack([ids, ids * 0, ids * 0 + 1], axis=1).astype(np.uint64)
    return meta, soa


def _worker(rank, world, port, n_games, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from c4a0_b200 import dist as D
    from c4a0_b200.nn import ConnectFourNet, ModelConfig

    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=2, n_policy_layers=2, n_value_layers=1))
    sent = D.broadcast_model(model)
    digest = float(sum(p.double().sum() for p in model.parameters()) + sum(b.double().sum() for b in model.buffers()))
    lo, hi = D.shard_range(n_games, rank, world)
    meta, soa = _fake_samples(lo, hi)
    gm, gs = D.gather_samples(meta, soa, device=torch.device("cpu"))
    ok = True
    if rank == 0:
        em, es = _fake_samples(0, n_games)
        ok = np.array_equal(gm, em) and all(
            np.array_equal(getattr(gs, f), getattr(es, f)) for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty")
        )
    else:
        ok = gm is None and gs is None
    q.put((rank, sent > 0, digest, (lo, hi), ok))
    torch.distributed.destroy_process_group()


def test_shard_range_partitions_exactly():
    from c4a0_b200.dist import shard_range

    for n in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_games = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_games, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, sent0, d0, rng0, ok0), (r1, sent1, d1, rng1, ok1) = got
    assert sent0 and sent1 and d0 == d1  # same weights everywhere after the broadcast
    assert rng0 == (0, 6) and rng1 == (6, 11)
    assert ok0 and ok1
