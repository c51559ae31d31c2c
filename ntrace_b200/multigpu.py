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


_comm_ready = False


def init_comm() -> bool:
    """Create the library's own NCCL communicator (nt_comm_init, C ABI) for the ranks of the default torch.distributed group: rank 0
    asks the library for the unique id, torch.distributed only carries those 128 bytes to the other ranks.  False when NCCL cannot
    be bound (the callers then fall back to torch's collectives)."""
    global _comm_ready
    if _comm_ready:
        return True
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        try:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        except capi.NtError:
            ok.zero_()
    dist.broadcast(ok, 0)
    if int(ok.item()) == 0:
        return False
    dist.broadcast(uid, 0)
    try:
        capi.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    except capi.NtError:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    _comm_ready = int(ok.item()) == 1
    return _comm_ready


def broadcast_bvh(src: int = 0, use_capi: bool = True) -> float:
    """Replicate the BVH resident on rank `src` to every rank (NCCL over NVLink).  Returns seconds (device time
    of the three broadcasts on this rank).  With use_capi the library's own nt_bvh_broadcast does it (the communicator of
    init_comm()); otherwise, or when NCCL could not be bound inside the library, torch.distributed broadcasts the three buffers."""
    import torch
    import torch.distributed as dist
    if use_capi and init_comm():
        return capi.bvh_broadcast(src)
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


# --------------------------------------------------------------------------------------------------
# Host placement: one process per GPU wants its pinned ray / result buffers on the NUMA node the GPU hangs off, otherwise
# every zero-copy read and DMA crosses the socket interconnect (measured at 8 GPUs: host-buffer Mrays/s per GPU drops
# to a third without this).
def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(pci_domain: int, pci_bus: int, pci_device: int, sysfs: str = "/sys") -> int:
    """NUMA node of a PCI device from sysfs, -1 when the platform does not say."""
    path = f"{sysfs}/bus/pci/devices/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0/numa_node"
    try:
        return int(open(path).read().strip())
    except (OSError, ValueError):
        return -1


def bind_host_to_gpu(local_rank: int, sysfs: str = "/sys"):
    """Restrict this process to the CPUs of the GPU's NUMA node (first-touch then places pinned memory there).
    Returns (previous affinity, node) so a caller can restore it; (None, -1) when nothing was changed."""
    import os
    import torch
    try:
        p = torch.cuda.get_device_properties(local_rank)
        node = gpu_numa_node(getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id, sysfs)
        if node < 0:
            return None, -1
        cpus = _parse_cpulist(open(f"{sysfs}/devices/system/node/node{node}/cpulist").read())
        prev = os.sched_getaffinity(0)
        want = cpus & prev
        if not want:
            return None, node
        os.sched_setaffinity(0, want)
        return prev, node
    except (OSError, AttributeError, RuntimeError, ValueError):
        return None, -1
