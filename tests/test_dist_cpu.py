"""world_size-2 gloo test of the N>1 host path: slab partition + optional final gather.  Each rank fills
its slab with the oracle (tests may use it) and the product's gather must reassemble the full batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wignerd_oracle as o
        from wignerd.jl_b200.shard import gather_slabs, shard_bounds
        betas = np.linspace(0, np.pi, n)
        lo, hi = shard_bounds(n, rank, world)
        local = torch.from_numpy(np.ascontiguousarray(o.wignerd_batch(1.5, betas[lo:hi])))
        full = gather_slabs(local, n, world)
        ref = torch.from_numpy(o.wignerd_batch(1.5, betas))
        ok = full.shape == ref.shape and bool(torch.equal(full, ref))
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [50, 37])
def test_gather_slabs_gloo_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1


def _worker_sym(rank, world, port, n, ret):
    """the bench's sharding: every rank owns a symmetric shard (front slab + mirror image); the ranks' element counts and
    checksums, reduced over gloo, must account for the whole batch exactly once"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wignerd_oracle as o
        from wignerd.jl_b200.shard import symmetric_shard
        betas = np.linspace(0, np.pi, n)
        (lo, hi), (blo, bhi) = symmetric_shard(n, rank, world)
        mine = np.concatenate([betas[lo:hi], betas[blo:bhi]])
        # the rank's own grid is symmetric about pi/2 (what the pairing of the contraction kernels needs)
        sym = len(mine) == 0 or float(np.abs(mine + mine[::-1] - np.pi).max()) < 2e-15
        d = o.wignerd_batch(2, mine) if len(mine) else np.zeros((0, 5, 5))
        t = torch.tensor([float(len(mine)), float(d.sum()), 1.0 if sym else 0.0], dtype=torch.float64)
        dist.all_reduce(t[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(t[2:], op=dist.ReduceOp.MIN)
        if rank == 0:
            ref = o.wignerd_batch(2, betas)
            ret.put(int(t[0].item() == n and abs(t[1].item() - ref.sum()) < 1e-9 and t[2].item() == 1.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 101])
def test_symmetric_shards_gloo_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker_sym, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1
