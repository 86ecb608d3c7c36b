"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, statistic merging, gradient all-reduce,
HAPPO order broadcast.  Envs need no collective; the oracle test `test_philox_source_is_deterministic_and_sharded`
shows that shards reproduce the unsharded run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from d2d_ppo_b200.algorithms import _dist
    try:
        assert _dist.active() and _dist.rank() == rank and _dist.world_size() == world
        # 1. env shards are contiguous, disjoint and cover the global index space
        off, cnt = _dist.shard(1_000_003)
        spans = [None] * world
        dist.all_gather_object(spans, (off, cnt))
        assert spans[0][0] == 0 and sum(c for _, c in spans) == 1_000_003
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert _dist.is_last_shard() == (rank == world - 1)
        # 2. normalisation statistics: merged (sum, sumsq) moments == moments of the concatenated data
        rng = np.random.default_rng(0)
        full = rng.normal(1.5, 2.0, (1000, 3))
        mine = full[off % 1000: off % 1000 + 0]  # (unused) keep shapes explicit below
        lo, n = _dist.shard(1000)
        part = torch.tensor(full[lo:lo + n])
        stats = torch.stack([part.sum(0), (part ** 2).sum(0)], dim=1)
        for ddof in (0, 1):
            mean, std = _dist.merge_moments(stats, n, ddof)
            assert np.allclose(mean.numpy(), full.mean(0), rtol=1e-12)
            assert np.allclose(std.numpy(), full.std(0, ddof=ddof), rtol=1e-10)
        # 3. gradient all-reduce: sum of per-shard gradients (each already scaled by 1 / global rows)
        g = torch.full((4, 7), float(rank + 1))
        _dist.all_reduce_sum_(g)
        assert torch.equal(g, torch.full((4, 7), float(sum(range(1, world + 1)))))
        # 4. every rank applies rank 0's HAPPO agent order
        order = np.random.default_rng(rank).permutation(6)
        got = _dist.broadcast_order(order, "cpu").tolist()
        assert got == np.random.default_rng(0).permutation(6).tolist()
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_single_process_defaults():
    from d2d_ppo_b200.algorithms import _dist
    assert not _dist.active() and _dist.rank() == 0 and _dist.world_size() == 1
    assert _dist.shard(10) == (0, 10) and _dist.is_last_shard()
    t = torch.ones(3)
    assert _dist.all_reduce_sum_(t) is t
    assert _dist.broadcast_order([2, 0, 1], "cpu").tolist() == [2, 0, 1]
