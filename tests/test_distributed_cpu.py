"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: round-robin batch sharding, the MIN/MAX range
all-reduce that precedes the EMA, and the eval-loss (sum, count) all-reduce."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from outeffhop_b200 import quantization as Q
    from outeffhop_b200.eval_harness import aggregate_perplexity, shard_indices
    # 1. range sync: global min/max over both ranks, bit-identical on every rank
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(1000, generator=g) * (1 + rank)
    mm = torch.stack((x.min(), x.max()))
    Q.set_range_sync(True)
    synced = Q._sync_minmax(mm)
    # 1b. a GROUP of independent taps (q / k / v projections): (n, 2) pairs cross the ranks in ONE all-reduce
    n0 = Q.RANGE_SYNC_COLLECTIVES
    grp = torch.stack([torch.stack(((x * s).min(), (x * s).max())) for s in (1.0, 0.5, 3.0)])
    grp_synced = Q._sync_minmax(grp)
    grp_calls = Q.RANGE_SYNC_COLLECTIVES - n0
    Q.set_range_sync(False)
    unsynced = Q._sync_minmax(mm)
    # 2. loss aggregation: every batch weighted equally across ranks
    idx = shard_indices(7, rank, world)
    losses = torch.tensor([1.0 + 0.1 * i for i in idx])
    ppl = aggregate_perplexity(losses.sum(), len(idx))
    q.put((rank, synced.tolist(), unsynced.tolist(), mm.tolist(), idx, ppl, grp.tolist(), grp_synced.tolist(), grp_calls))
    dist.destroy_process_group()


def test_range_sync_and_loss_aggregation_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, u0, m0, i0, p0, g0, gs0, c0), (r1, s1, u1, m1, i1, p1, g1, gs1, c1) = res
    assert c0 == c1 == 1                                               # three taps, one collective
    assert gs0 == gs1 == [[min(a[0], b[0]), max(a[1], b[1])] for a, b in zip(g0, g1)]
    assert s0 == s1 == [min(m0[0], m1[0]), max(m0[1], m1[1])]          # exact, identical on both ranks
    assert u0 == m0 and u1 == m1                                       # disabled: per-process, like the reference
    assert i0 == [0, 2, 4, 6] and i1 == [1, 3, 5]
    import math
    expect = math.exp(sum(1.0 + 0.1 * i for i in range(7)) / 7)
    assert abs(p0 - expect) < 1e-6 and abs(p1 - expect) < 1e-6
