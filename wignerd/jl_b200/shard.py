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


def shard_bounds_native(n: int, rank: int, world: int, align: int = 16) -> tuple[int, int]:
    """the same arithmetic through the C ABI (wd_shard_bounds): what a non-Python host uses"""
    import ctypes
    from .api import load_library, _check
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    _check(load_library().wd_shard_bounds(int(n), int(rank), int(world), int(align), ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def symmetric_shard(n: int, rank: int, world: int, align: int = 16) -> tuple[tuple[int, int], tuple[int, int]]:
    """A shard that is closed under b -> n-1-b: the rank's slab [lo, hi) of the FIRST half of the batch plus its mirror image
    [n-hi, n-lo) in the second half.  On angle grids that are symmetric about pi/2 (linspace(0, pi, n)) the rank's own
    angle list -- front slab followed by back slab -- is then symmetric too, which is what the contraction kernel's
    beta <-> pi - beta pairing needs.  For odd n the middle element goes to the last rank's front slab."""
    half = (n + 1) // 2
    lo, hi = shard_bounds(half, rank, world, align)
    blo, bhi = n - hi, n - lo
    if n % 2 == 1 and hi == half:         # the middle element is its own mirror: keep it once
        blo += 1
    return (lo, hi), (min(blo, bhi), bhi)


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
