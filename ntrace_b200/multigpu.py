"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

The path shards over rays with no data-path collective: the BVH is replicated once per (re)build with a
broadcast of the three CudaBVH buffers, and GPU ``g`` of ``G`` traces the contiguous slot range
``[g*ceil(N/G), min(N, (g+1)*ceil(N/G)))`` of a RayBuffer (contiguity keeps the PixelTable / Morton coherence
inside a slice).  Results stay in place on each rank; ``gather_results`` is provided for callers that need the
whole batch on every rank (not on the timed path in the reference's accounting, which times kernels only).

The reference has no multi-GPU support (single CUDA context, src/framework/gpu/CudaModule.hpp:92-97); this
module is new functionality around the same per-GPU trace call.
"""
from __future__ import annotations

import numpy as np

from . import capi


def slice_for_rank(n: int, rank: int, world: int):
    """Contiguous slot range of a batch of n rays owned by `rank` -> (lo, hi)."""
    if world <= 0 or rank < 0 or rank >= world:
        raise ValueError("bad rank/world")
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, (rank + 1) * per)
    return lo, hi


class _DevicePtr:
    """Exposes a raw device allocation of the library to torch (zero copy) via __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _bvh_tensors():
    import torch
    (nb, wb, ib), layout = capi.bvh_sizes()
    ptrs = capi.bvh_device_ptrs()
    return [torch.as_tensor(_DevicePtr(p, n), device="cuda") for p, n in zip(ptrs, (nb, wb, ib))], layout


def broadcast_meta(meta, src: int):
    """Broadcast [layout, nodeBytes, woopBytes, idxBytes] (int64 tensor on the collective's device)."""
    import torch.distributed as dist
    dist.broadcast(meta, src)
    return [int(x) for x in meta.tolist()]


def broadcast_bvh(src: int = 0) -> float:
    """Replicate the BVH resident on rank `src` to every rank (NCCL over NVLink).  Returns seconds (device time
    of the three broadcasts on this rank)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = torch.zeros(4, dtype=torch.int64, device="cuda")
    if rank == src:
        (nb, wb, ib), layout = capi.bvh_sizes()
        meta.copy_(torch.tensor([layout, nb, wb, ib], dtype=torch.int64))
    layout, nb, wb, ib = broadcast_meta(meta, src)
    if rank != src:
        capi.bvh_alloc(layout, nb, wb, ib)
    capi.synchronize()
    tensors, _ = _bvh_tensors()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in tensors:
        dist.broadcast(t, src)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def gather_results(local, n: int, rank: int, world: int):
    """All-gather the per-rank result slices ([hi-lo, 4] int32 tensors) into the full [n, 4] batch."""
    import torch
    import torch.distributed as dist
    per = -(-n // world) if n > 0 else 0
    pad = torch.zeros((per, 4), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat(parts, 0)[:n]


def trace_batch_sharded(tracer, rays, rank: int, world: int) -> float:
    """Trace this rank's slice of a host.RayBuffer in place; returns kernel seconds on this rank."""
    lo, hi = slice_for_rank(rays.getSize(), rank, world)
    if hi <= lo:
        return 0.0
    from . import host
    host._sync()
    return capi.trace_batch(rays.getRayBuffer()[lo:hi], rays.getResultBuffer()[lo:hi], hi - lo, rays.getNeedClosestHit())
