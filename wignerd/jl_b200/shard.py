"""Multi-GPU plumbing: the (j, beta) batch shards with NO inter-GPU traffic on the hot path
 -- every output matrix depends only on its own angle and the small per-j table,
which each rank builds for itself.  One process per GPU; contiguous slabs of the angle batch per rank.
The only collective is the OPTIONAL final gather of the per-rank output slabs (NCCL over NVLink on
GPUs, gloo in the CPU tests), which is never part of the timed hot path.
"""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int, align: int = 16) -> tuple[int, int]:
    """[lo, hi) of the contiguous slab of ``n`` batch items owned by ``rank``.  Slab starts are aligned to
    ``align`` items (the kernel's angle-chunk size) so no chunk straddles two ranks; sizes differ by at
    most ``align``; the slabs partition [0, n) exactly."""
    if world < 1 or not (0 <= rank < world):
        raise AssertionError("bad rank/world")
    if n < 0:
        raise AssertionError("negative batch size")
    nchunks = (n + align - 1) // align
    base, rem = divmod(nchunks, world)
    lo_c = rank * base + min(rank, rem)
    hi_c = lo_c + base + (1 if rank < rem else 0)
    return min(lo_c * align, n), min(hi_c * align, n)


def gather_slabs(local, n: int, world: int, group=None, align: int = 16):
    """Optional final gather: every rank contributes its slab ``local`` (first dim = its share of the
    batch) and receives the full batch.  Uses torch.distributed.all_gather on padded slabs (works with
    NCCL and gloo).  Returns a tensor whose first dim is n."""
    import torch
    import torch.distributed as dist

    sizes = [shard_bounds(n, r, world, align) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
