"""ctypes binding of ``libntrace_b200.so`` (C ABI: ``include/ntrace_b200.h``).

This is the only way Python reaches the CUDA kernels.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is present, every entry point raises :class:`NtError`.

Buffers may be numpy arrays (host) or torch tensors (host or CUDA); the library detects the
address space itself, exactly as the C ABI documents.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NTRACE_B200_LIB") or os.path.join(_HERE, "libntrace_b200.so")      # override: development builds only

LAYOUT_COMPACT = 4
LAYOUT_COMPACT2 = 5
BUILDER_LBVH = 0
BUILDER_HLBVH = 1

# every symbol include/ntrace_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "nt_init", "nt_shutdown", "nt_last_error", "nt_launch_count",
    "nt_mem_alloc", "nt_mem_free", "nt_mem_alloc_host", "nt_mem_free_host", "nt_memcpy", "nt_memset",
    "nt_event_record", "nt_event_elapsed", "nt_set_deferred", "nt_synchronize",
    "nt_set_kernel", "nt_desired_layout", "nt_kernel_config",
    "nt_bvh_upload", "nt_bvh_alloc", "nt_bvh_build", "nt_bvh_set_collapse", "nt_bvh_set_build_layout", "nt_bvh_convert", "nt_bvh_sizes", "nt_bvh_download",
    "nt_bvh_device_ptrs", "nt_bvh_build_debug", "nt_bvh_wide4_convert_host", "nt_bvh_wide4_download", "nt_raygen_set_order", "nt_trace_batches", "nt_bvh_generation", "nt_bvh_sah", "nt_hash_buffer", "nt_comm_unique_id", "nt_comm_init", "nt_comm_destroy", "nt_comm_allreduce", "nt_bvh_broadcast",
    "nt_trace_batch", "nt_trace_batch_async", "nt_trace_wait", "nt_raygen_primary", "nt_raygen_ao", "nt_raygen_shadow", "nt_ray_sort", "nt_count_hits", "nt_tri_normals",
]


class NtError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA extension; fail loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NtError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "ntrace_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        l.nt_last_error.restype = C.c_char_p
        l.nt_launch_count.restype = C.c_int64
        l.nt_shutdown.restype = None
        _lib = l
    return _lib


def _check(rc: int):
    if rc != 0:
        raise NtError(lib().nt_last_error().decode("utf-8", "replace"))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def ptr(x, dtype=None, min_bytes: int = 0):
    """Raw address of a numpy array / torch tensor (contiguous), or None."""
    if x is None:
        return None
    if _is_torch(x):
        if not x.is_contiguous():
            raise NtError("tensor must be contiguous")
        nbytes = x.numel() * x.element_size()
        addr = x.data_ptr()
    else:
        if not isinstance(x, np.ndarray) or not x.flags["C_CONTIGUOUS"]:
            raise NtError("buffer must be a C-contiguous numpy array or a torch tensor")
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise NtError(f"expected dtype {np.dtype(dtype)}, got {x.dtype}")
        nbytes = x.nbytes
        addr = x.ctypes.data
    if nbytes < min_bytes:
        raise NtError(f"buffer too small: {nbytes} < {min_bytes} bytes")
    return C.c_void_p(addr)


def _nbytes(x) -> int:
    return x.numel() * x.element_size() if _is_torch(x) else x.nbytes


# ---- lifetime ------------------------------------------------------------------------------------
def init(device: int = 0):
    _check(lib().nt_init(C.c_int(device)))


def shutdown():
    lib().nt_shutdown()


def launch_count() -> int:
    return int(lib().nt_launch_count())


def event_record(slot: int):
    _check(lib().nt_event_record(C.c_int(slot)))


def event_elapsed(a: int, b: int) -> float:
    sec = C.c_float(0.0)
    _check(lib().nt_event_elapsed(C.c_int(a), C.c_int(b), C.byref(sec)))
    return float(sec.value)


def set_deferred(enabled):
    """False / 0: synchronous calls; True / 1: launches are queued on the main stream; 2: queued launches alternate between two
    kernel streams so that consecutive batches overlap at their tails (batches in flight need distinct result buffers)."""
    _check(lib().nt_set_deferred(C.c_int(int(enabled))))


def synchronize():
    _check(lib().nt_synchronize())


# ---- kernel selection -----------------------------------------------------------------------------
def set_kernel(name: str):
    _check(lib().nt_set_kernel(name.encode()))


def desired_layout() -> int:
    return int(lib().nt_desired_layout())


def kernel_config() -> dict:
    out = (C.c_int32 * 4)()
    _check(lib().nt_kernel_config(out))
    return dict(bvhLayout=out[0], blockWidth=out[1], blockHeight=out[2], usePersistentThreads=out[3])


# ---- BVH -----------------------------------------------------------------------------------------
def bvh_upload(layout: int, nodes, woop, tri_index):
    _check(lib().nt_bvh_upload(C.c_int(layout), ptr(nodes), C.c_size_t(_nbytes(nodes)), ptr(woop), C.c_size_t(_nbytes(woop)),
                               ptr(tri_index), C.c_size_t(_nbytes(tri_index))))


def bvh_alloc(layout: int, node_bytes: int, woop_bytes: int, idx_bytes: int):
    _check(lib().nt_bvh_alloc(C.c_int(layout), C.c_size_t(node_bytes), C.c_size_t(woop_bytes), C.c_size_t(idx_bytes)))


def bvh_build(builder: int, verts, tris, bbox_lo, bbox_hi, hlbvh_bits=4, leaf_size=8, epsilon=0.001) -> float:
    """GPU LBVH/HLBVH build -> GPU seconds (CUDA events around the whole pipeline)."""
    nv = _nbytes(verts) // 12
    nt = _nbytes(tris) // 12
    lo = (C.c_float * 3)(*[float(v) for v in bbox_lo])
    hi = (C.c_float * 3)(*[float(v) for v in bbox_hi])
    sec = C.c_float(0.0)
    _check(lib().nt_bvh_build(C.c_int(builder), ptr(verts, np.float32), C.c_int(nv), ptr(tris, np.int32), C.c_int(nt), lo, hi,
                              C.c_int(hlbvh_bits), C.c_int(leaf_size), C.c_float(epsilon), C.byref(sec)))
    return float(sec.value)


def bvh_set_build_layout(layout: int):
    """Layout nt_bvh_build emits: 4 Compact (default) or 5 Compact2 (needed beyond 31 M inner nodes)."""
    _check(lib().nt_bvh_set_build_layout(C.c_int(layout)))


def bvh_convert(layout: int):
    """Make the resident BVH Compact (4) / Compact2 (5) in place."""
    _check(lib().nt_bvh_convert(C.c_int(layout)))


def bvh_set_collapse(mode: int, max_leaf: int = 0):
    _check(lib().nt_bvh_set_collapse(C.c_int(mode), C.c_int(max_leaf)))


def bvh_sah() -> dict:
    """SAH (BVHNode.cpp:79-94 formula) and node / leaf / triangle counts of the resident BVH, computed on the device."""
    sah = C.c_double(0.0)
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _check(lib().nt_bvh_sah(C.byref(sah), C.byref(a), C.byref(b), C.byref(c)))
    return dict(sah=float(sah.value), num_inner=int(a.value), num_leaf=int(b.value), num_tris=int(c.value))


def hash_buffer(data) -> int:
    """FW::hashBuffer (Hash.cpp:33-75) of a bytes-like object / numpy array; host-only."""
    if isinstance(data, (bytes, bytearray)):
        data = np.frombuffer(bytes(data), dtype=np.uint8)
    data = np.ascontiguousarray(data)
    out = C.c_uint32(0)
    _check(lib().nt_hash_buffer(C.c_void_p(data.ctypes.data) if data.nbytes else None, C.c_size_t(data.nbytes), C.byref(out)))
    return int(out.value)


def comm_unique_id() -> bytes:
    buf = (C.c_char * 128)()
    _check(lib().nt_comm_unique_id(buf))
    return bytes(buf)


def comm_init(num_ranks: int, rank: int, unique_id: bytes):
    _check(lib().nt_comm_init(C.c_int(num_ranks), C.c_int(rank), C.c_char_p(unique_id)))


def comm_destroy():
    _check(lib().nt_comm_destroy())


def comm_allreduce(values, op: str = "max"):
    a = np.ascontiguousarray(values, dtype=np.float64).copy()
    _check(lib().nt_comm_allreduce(ptr(a), C.c_int(len(a)), C.c_int(0 if op == "sum" else 1)))
    return a


def bvh_broadcast(root: int = 0) -> float:
    """Replicate rank `root`'s resident BVH into every rank's library (NCCL, inside the C ABI) -> seconds of the three broadcasts."""
    sec = C.c_float(0.0)
    _check(lib().nt_bvh_broadcast(C.c_int(root), C.byref(sec)))
    return float(sec.value)


def bvh_generation() -> int:
    """Generation counter of the library's ONE resident BVH (0 = none); bumped by every upload / alloc / build / convert."""
    out = C.c_uint64(0)
    _check(lib().nt_bvh_generation(C.byref(out)))
    return int(out.value)


def bvh_sizes():
    sizes = (C.c_size_t * 3)()
    layout = C.c_int(0)
    _check(lib().nt_bvh_sizes(sizes, C.byref(layout)))
    return (int(sizes[0]), int(sizes[1]), int(sizes[2])), int(layout.value)


def bvh_download():
    (nb, wb, ib), layout = bvh_sizes()
    nodes = np.zeros(nb // 4, dtype=np.int32)
    woop = np.zeros(wb // 4, dtype=np.int32)
    idx = np.zeros(ib // 4, dtype=np.int32)
    _check(lib().nt_bvh_download(ptr(nodes), ptr(woop), ptr(idx)))
    return nodes, woop, idx, layout


def bvh_wide4_convert_host(layout: int, nodes, woop_bytes: int):
    """Wide4 node array (uint32, 16 words per node) of a Compact / Compact2 node buffer; host-only -> (wide_nodes, max_depth)."""
    size = C.c_size_t(0)
    depth = C.c_int(0)
    _check(lib().nt_bvh_wide4_convert_host(C.c_int(layout), ptr(nodes), C.c_size_t(_nbytes(nodes)), C.c_size_t(woop_bytes), None, C.c_size_t(0),
                                           C.byref(size), C.byref(depth)))
    out = np.zeros(size.value // 4, dtype=np.uint32)
    _check(lib().nt_bvh_wide4_convert_host(C.c_int(layout), ptr(nodes), C.c_size_t(_nbytes(nodes)), C.c_size_t(woop_bytes), ptr(out), C.c_size_t(out.nbytes),
                                           C.byref(size), C.byref(depth)))
    return out, int(depth.value)


def trace_batches(rays_list, results_list, counts, need_closest: bool) -> float:
    """nt_trace_batches: several device-resident batches (torch tensors) in one persistent launch -> GPU seconds (0 when deferred)."""
    n = len(rays_list)
    rp = (C.c_void_p * n)(*[ptr(r) for r in rays_list])
    op = (C.c_void_p * n)(*[ptr(r) for r in results_list])
    cnt = (C.c_int32 * n)(*[int(c) for c in counts])
    sec = C.c_float(0.0)
    _check(lib().nt_trace_batches(C.c_int(n), rp, op, cnt, C.c_int(1 if need_closest else 0), C.byref(sec)))
    return float(sec.value)


def raygen_set_order(mode: int):
    """0: nt_raygen_ao writes the reference's slot order; 1: direction-coherent order inside tiles of <= 1024 rays."""
    _check(lib().nt_raygen_set_order(C.c_int(mode)))


def bvh_wide4_download():
    """Wide4 node array of the resident BVH as derived on the device -> (wide_nodes uint32, max_depth)."""
    size = C.c_size_t(0)
    depth = C.c_int(0)
    _check(lib().nt_bvh_wide4_download(None, C.c_size_t(0), C.byref(size), C.byref(depth)))
    out = np.zeros(size.value // 4, dtype=np.uint32)
    _check(lib().nt_bvh_wide4_download(ptr(out), C.c_size_t(out.nbytes), C.byref(size), C.byref(depth)))
    return out, int(depth.value)


def bvh_device_ptrs():
    p = (C.c_void_p * 3)()
    _check(lib().nt_bvh_device_ptrs(p))
    return [int(p[i] or 0) for i in range(3)]


def bvh_build_debug(num_tris: int):
    keys = np.zeros(num_tris, dtype=np.uint32)
    idx = np.zeros(num_tris, dtype=np.int32)
    _check(lib().nt_bvh_build_debug(ptr(keys), ptr(idx), C.c_int(num_tris)))
    return keys, idx


# ---- trace / raygen --------------------------------------------------------------------------------
def trace_batch(rays, results, num_rays: int, need_closest_hit: bool) -> float:
    """CudaBVHTracer::traceBatch -> kernel seconds (CUDA events around the launch only)."""
    sec = C.c_float(0.0)
    _check(lib().nt_trace_batch(ptr(rays, np.float32, num_rays * 32), ptr(results, np.int32, num_rays * 16), C.c_int(num_rays),
                                C.c_int(1 if need_closest_hit else 0), C.byref(sec)))
    return float(sec.value)


ASYNC_SLOTS = 4


def trace_batch_async(rays, results, num_rays: int, need_closest_hit: bool, slot: int):
    """Submit a batch into `slot` (device or pinned host buffers); collect with trace_wait(slot)."""
    _check(lib().nt_trace_batch_async(ptr(rays, np.float32, num_rays * 32), ptr(results, np.int32, num_rays * 16), C.c_int(num_rays),
                                      C.c_int(1 if need_closest_hit else 0), C.c_int(slot)))


def trace_wait(slot: int) -> float:
    sec = C.c_float(0.0)
    _check(lib().nt_trace_wait(C.c_int(slot), C.byref(sec)))
    return float(sec.value)


def raygen_primary(rays, id_to_slot, slot_to_id, origin, nscreen_to_world, w: int, h: int, max_dist: float, seed: int = 0):
    o = (C.c_float * 3)(*[float(v) for v in origin])
    m = (C.c_float * 16)(*[float(v) for v in np.asarray(nscreen_to_world, dtype=np.float32).reshape(-1)])
    _check(lib().nt_raygen_primary(ptr(rays, np.float32, w * h * 32), ptr(id_to_slot, np.int32), ptr(slot_to_id, np.int32), o, m,
                                   C.c_int(w), C.c_int(h), C.c_float(max_dist), C.c_uint32(seed)))


def raygen_ao(out_rays, out_id_to_slot, out_slot_to_id, in_rays, in_results, tri_normals, first: int, count: int, samples: int,
              max_dist: float, seed: int):
    _check(lib().nt_raygen_ao(ptr(out_rays, np.float32, count * samples * 32), ptr(out_id_to_slot, np.int32), ptr(out_slot_to_id, np.int32),
                              ptr(in_rays, np.float32), ptr(in_results, np.int32), ptr(tri_normals, np.float32),
                              C.c_int(first), C.c_int(count), C.c_int(samples), C.c_float(max_dist), C.c_uint32(seed)))


def raygen_shadow(out_rays, out_id_to_slot, out_slot_to_id, in_rays, in_results, first: int, count: int, samples: int,
                  light_pos, light_radius: float, seed: int):
    lp = (C.c_float * 3)(*[float(v) for v in light_pos])
    _check(lib().nt_raygen_shadow(ptr(out_rays, np.float32, count * samples * 32), ptr(out_id_to_slot, np.int32), ptr(out_slot_to_id, np.int32),
                                  ptr(in_rays, np.float32), ptr(in_results, np.int32), C.c_int(first), C.c_int(count), C.c_int(samples),
                                  lp, C.c_float(light_radius), C.c_uint32(seed)))


def ray_sort(rays, id_to_slot, slot_to_id, num_rays: int):
    """RayBuffer::mortonSort, in place."""
    _check(lib().nt_ray_sort(ptr(rays, np.float32, num_rays * 32), ptr(id_to_slot, np.int32, num_rays * 4), ptr(slot_to_id, np.int32, num_rays * 4),
                             C.c_int(num_rays)))


def count_hits(results, num_rays: int) -> int:
    out = C.c_int(0)
    _check(lib().nt_count_hits(ptr(results, np.int32, num_rays * 16), C.c_int(num_rays), C.byref(out)))
    return int(out.value)


def tri_normals(verts, tris, out):
    nv = _nbytes(verts) // 12
    nt = _nbytes(tris) // 12
    _check(lib().nt_tri_normals(ptr(verts, np.float32), C.c_int(nv), ptr(tris, np.int32), C.c_int(nt), ptr(out, np.float32, nt * 12)))
