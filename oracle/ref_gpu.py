"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/_ref/libref_<kernel>.so``: the reference's OWN traversal kernels
(``src/rt/kernels/fermi_speculative_while_while.cu``, ``kepler_dynamic_fetch.cu``) compiled for sm_100a from where
they lie under ``/root/reference`` (``make -C oracle ref_gpu``; shims and the two syntactic patches are listed in
``oracle/Makefile`` / ``oracle/ref_shim/ref_gpu_trace.cu``).  They run on the GPU box only.

Used by ``tests/test_gpu_reference_kernels.py`` (the B200 kernel against the recompiled reference kernels on the same BVH
and rays) and by ``bench.py`` to report the recompiled reference's Mrays/s beside the B200 kernel's.  Nothing under
``ntrace_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS = ("fermi_speculative_while_while", "kepler_dynamic_fetch")
_libs = {}


def lib_path(kernel: str) -> str:
    return os.path.join(_HERE, "_ref", f"libref_{kernel}.so")


def available(kernel: str = KERNELS[0]) -> bool:
    return os.path.exists(lib_path(kernel))


def _lib(kernel: str):
    if kernel not in _libs:
        if kernel not in KERNELS or not available(kernel):
            raise RuntimeError(f"oracle/_ref/libref_{kernel}.so not built (needs /root/reference; run `make -C oracle ref_gpu`)")
        l = C.CDLL(lib_path(kernel), mode=os.RTLD_LOCAL)
        l.ref_gpu_error.restype = C.c_char_p
        _libs[kernel] = l
    return _libs[kernel]


def _dp(t):
    return C.c_void_p(int(t.data_ptr()))


def _raygen_lib(ieee: bool):
    name = "raygen_ieee" if ieee else "raygen_fast"
    if name not in _libs:
        path = os.path.join(_HERE, "_ref", f"libref_{name}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built (run `make -C oracle ref_gpu`)")
        l = C.CDLL(path, mode=os.RTLD_LOCAL)
        l.ref_gpu_error.restype = C.c_char_p
        _libs[name] = l
    return _libs[name]


def raygen_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"libref_raygen_{v}.so")) for v in ("fast", "ieee"))


def _ck(l, rc):
    if rc != 0:
        raise RuntimeError("reference kernel failed: " + l.ref_gpu_error().decode())


def raygen_primary(rays, id_to_slot, slot_to_id, index_to_pixel, origin, n2w, w, h, max_dist, seed=0, ieee=True):
    """rayGenPrimaryKernel (RayGenKernels.cu:77-125) on torch CUDA tensors; n2w = 4x4 row-major numpy."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _raygen_lib(ieee)
    o = (C.c_float * 3)(*[float(v) for v in origin])
    m = (C.c_float * 16)(*[float(v) for v in np.asarray(n2w, dtype=np.float32).reshape(-1)])
    _ck(l, l.ref_raygen_primary(_dp(rays), _dp(id_to_slot), _dp(slot_to_id), _dp(index_to_pixel), o, m, C.c_int(w), C.c_int(h), C.c_float(max_dist), C.c_uint(seed)))


def raygen_ao(out_rays, out_i2s, out_s2i, in_rays, in_results, normals, first, count, samples, max_dist, seed, ieee=True):
    """rayGenAOKernel (RayGenKernels.cu:129-236)."""
    import torch
    torch.cuda.synchronize()
    l = _raygen_lib(ieee)
    _ck(l, l.ref_raygen_ao(_dp(out_rays), _dp(out_i2s), _dp(out_s2i), _dp(in_rays), _dp(in_results), _dp(normals), C.c_int(first), C.c_int(count),
                           C.c_int(samples), C.c_float(max_dist), C.c_uint(seed)))


def raygen_shadow(out_rays, out_i2s, out_s2i, in_rays, in_results, first, count, samples, light_pos, light_radius, seed, ieee=True):
    """rayGenShadowKernel (RayGenKernels.cu:240-302)."""
    import torch
    torch.cuda.synchronize()
    l = _raygen_lib(ieee)
    lp = (C.c_float * 3)(*[float(v) for v in light_pos])
    _ck(l, l.ref_raygen_shadow(_dp(out_rays), _dp(out_i2s), _dp(out_s2i), _dp(in_rays), _dp(in_results), C.c_int(first), C.c_int(count), C.c_int(samples),
                               lp, C.c_float(light_radius), C.c_uint(seed)))


def _named_lib(name: str):
    if name not in _libs:
        path = os.path.join(_HERE, "_ref", f"libref_{name}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built (run `make -C oracle ref_gpu`)")
        l = C.CDLL(path, mode=os.RTLD_LOCAL)
        l.ref_gpu_error.restype = C.c_char_p
        _libs[name] = l
    return _libs[name]


def hlbvh_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"libref_hlbvh_{v}.so")) for v in ("fast", "ieee"))


def raybuf_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", f"libref_raybuf_{v}.so")) for v in ("fast", "ieee"))


def ray_aabb(rays, ieee=True):
    """findAABBKernel (RayBufferKernels.cu:70-136) -> (lo[3], hi[3]) float32."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _named_lib("raybuf_ieee" if ieee else "raybuf_fast")
    out = (C.c_float * 6)()
    _ck(l, l.ref_ray_aabb(_dp(rays), C.c_int(int(rays.shape[0])), out))
    a = np.array(list(out), dtype=np.float32)
    return a[:3], a[3:]


def ray_keys(rays, lo, hi, ieee=True):
    """genMortonKeysKernel (RayBufferKernels.cu:140-175) -> uint32 [N, 6] 192-bit keys (hash[0] least significant)."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _named_lib("raybuf_ieee" if ieee else "raybuf_fast")
    n = int(rays.shape[0])
    raw = np.zeros((n, 7), dtype=np.uint32)
    flo = (C.c_float * 3)(*[float(v) for v in lo]); fhi = (C.c_float * 3)(*[float(v) for v in hi])
    _ck(l, l.ref_ray_keys(_dp(rays), C.c_int(n), flo, fhi, raw.ctypes.data_as(C.c_void_p)))
    assert (raw[:, 0] == np.arange(n, dtype=np.uint32)).all()
    return raw[:, 1:].copy()


def morton(verts, tris, lo, hi, ieee=True):
    """calcMorton (emitTreeKernel.cu:647-691) on torch CUDA tensors -> uint32 codes per triangle (numpy)."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _named_lib("hlbvh_ieee" if ieee else "hlbvh_fast")
    n = int(tris.shape[0])
    out = np.zeros(n, dtype=np.uint32)
    flo = (C.c_float * 3)(*[float(v) for v in lo]); fhi = (C.c_float * 3)(*[float(v) for v in hi])
    _ck(l, l.ref_morton(_dp(verts), _dp(tris), C.c_int(n), flo, fhi, out.ctypes.data_as(C.c_void_p)))
    return out


def lbvh_build(verts, tris, lo, hi, leaf_size=8, epsilon=0.001, ieee=True):
    """HLBVHBuilder::buildLBVH with the reference's own kernels -> dict(nodes, woop, tri_index, sorted_keys, sorted_idx,
    num_nodes, num_leaves, levels); node / leaf numbering depends on the order of atomics, compare in canonical form."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _named_lib("hlbvh_ieee" if ieee else "hlbvh_fast")
    n = int(tris.shape[0])
    flo = (C.c_float * 3)(*[float(v) for v in lo]); fhi = (C.c_float * 3)(*[float(v) for v in hi])
    info = (C.c_int * 3)()
    _ck(l, l.ref_lbvh_build(_dp(verts), _dp(tris), C.c_int(n), flo, fhi, C.c_int(leaf_size), C.c_float(epsilon), info))
    sizes = (C.c_longlong * 3)()
    _ck(l, l.ref_build_sizes(sizes))
    nodes = np.zeros(sizes[0] // 4, dtype=np.int32); woop = np.zeros(sizes[1] // 4, dtype=np.int32); idx = np.zeros(sizes[2] // 4, dtype=np.int32)
    keys = np.zeros(n, dtype=np.uint32); order = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _ck(l, l.ref_build_download(p(nodes), p(woop), p(idx), p(keys), p(order)))
    l.ref_build_gpu_ms.restype = C.c_float
    return dict(nodes=nodes, woop=woop, tri_index=idx, sorted_keys=keys, sorted_idx=order, num_nodes=info[0], num_leaves=info[1], levels=info[2],
                gpu_ms=float(l.ref_build_gpu_ms()))


def hlbvh_build(verts, tris, lo, hi, hlbvh_bits=4, leaf_size=8, epsilon=0.001, ieee=True):
    """HLBVHBuilder::buildHLBVH with the reference's own kernels (clusters + binned-SAH top level + LBVH below)."""
    import numpy as np
    import torch
    torch.cuda.synchronize()
    l = _named_lib("hlbvh_ieee" if ieee else "hlbvh_fast")
    n = int(tris.shape[0])
    flo = (C.c_float * 3)(*[float(v) for v in lo]); fhi = (C.c_float * 3)(*[float(v) for v in hi])
    info = (C.c_int * 6)()
    _ck(l, l.ref_hlbvh_build(_dp(verts), _dp(tris), C.c_int(n), flo, fhi, C.c_int(hlbvh_bits), C.c_int(leaf_size), C.c_float(epsilon), info))
    sizes = (C.c_longlong * 3)()
    _ck(l, l.ref_build_sizes(sizes))
    nodes = np.zeros(sizes[0] // 4, dtype=np.int32); woop = np.zeros(sizes[1] // 4, dtype=np.int32); idx = np.zeros(sizes[2] // 4, dtype=np.int32)
    keys = np.zeros(n, dtype=np.uint32); order = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _ck(l, l.ref_build_download(p(nodes), p(woop), p(idx), p(keys), p(order)))
    l.ref_build_gpu_ms.restype = C.c_float
    return dict(nodes=nodes, woop=woop, tri_index=idx, sorted_keys=keys, sorted_idx=order, num_nodes=info[0], num_leaves=info[1], levels=info[2],
                num_clusters=info[3], num_top_nodes=info[4], num_lbvh_roots=info[5], gpu_ms=float(l.ref_build_gpu_ms()))


def trace(kernel: str, rays, results, nodes, woop, tri_index, any_hit: bool = False, desired_warps: int = 0, repeats: int = 1):
    """Launch the reference kernel on torch CUDA tensors (rays [N,8] f32, results [N,4] i32, Compact / Compact2 BVH buffers).
    desired_warps = 0 keeps the reference's own launch size (CudaBVHTracer.cpp:151-156: one warp per 32 rays, or the hard-coded
    720 warps for persistent kernels).  Returns (best milliseconds over `repeats`, KernelConfig as a dict)."""
    import torch
    torch.cuda.synchronize()
    ms = C.c_float(0.0)
    cfg = (C.c_int * 4)()
    rc = _lib(kernel).ref_gpu_trace(_dp(rays), _dp(results), C.c_int(int(rays.shape[0])), C.c_int(1 if any_hit else 0), _dp(nodes), _dp(woop), _dp(tri_index),
                                    C.c_int(desired_warps), C.c_int(repeats), C.byref(ms), cfg)
    if rc != 0:
        raise RuntimeError("reference kernel failed: " + _lib(kernel).ref_gpu_error().decode())
    return float(ms.value), dict(bvhLayout=cfg[0], blockWidth=cfg[1], blockHeight=cfg[2], usePersistentThreads=cfg[3])
