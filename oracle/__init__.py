"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/liborc.so``, the CPU restatement of NTrace's tracing / build path
(see the headers of ``orc_*.hpp`` for the reference file:line each function follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``ntrace_b200/`` does.

Parity status: the reference ships no tests with golden values, so the CPU part of this
restatement (builders, SAH metric, createCompact/Woop, both tracers, Intersect primitives, pixel
table) is PINNED bit-for-bit against the reference's own sources compiled unmodified
(``oracle/ref.py``, ``oracle/_ref/libref.so``; ``tests/test_reference_pin.py``).  The restated
HLBVH builder and ray generators mirror device code: they are pinned on the GPU box against the
reference's own kernels compiled for sm_100a (``oracle/ref_gpu.py``,
``tests/test_gpu_reference_kernels.py``) and cross-validated in ``tests/test_oracle_*.py`` here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None

BUILDER_SAH = 0
BUILDER_SPLIT = 1


def build(force: bool = False) -> str:
    """Compile liborc.so with the committed Makefile (g++, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_bvh_build.restype = C.c_void_p
        _lib.orc_lbvh_build.restype = C.c_void_p
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_count_hits.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ----------------------------------------------------------------------------------------------
# CPU BVH (SAHBVHBuilder / SplitBVHBuilder) + BVH::trace + createCompact
# ----------------------------------------------------------------------------------------------
@dataclass
class TreeStats:
    sah: float
    num_inner: int
    num_leaf: int
    num_tris: int
    max_depth: int
    duplicates: int
    build_seconds: float


class CpuBVH:
    """Pointer-tree BVH built by the restated SAHBVHBuilder / SplitBVHBuilder."""

    def __init__(self, verts, tris, builder=BUILDER_SPLIT, min_leaf=1, max_leaf=1, split_alpha=1.0e-5):
        self.verts = _f32(verts).reshape(-1, 3)
        self.tris = _i32(tris).reshape(-1, 3)
        self._h = C.c_void_p(lib().orc_bvh_build(_p(self.verts), C.c_int(len(self.verts)), _p(self.tris), C.c_int(len(self.tris)),
                                                 C.c_int(builder), C.c_int(min_leaf), C.c_int(max_leaf), C.c_float(split_alpha)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bvh_free(self._h)
            self._h = None

    def stats(self) -> TreeStats:
        out = np.zeros(7, dtype=np.float64)
        lib().orc_bvh_stats(self._h, _p(out))
        return TreeStats(float(out[0]), int(out[1]), int(out[2]), int(out[3]), int(out[4]), int(out[5]), float(out[6]))

    def trace(self, rays, need_closest=True, counters=False, nthreads=0):
        """BVH::trace (Moller-Trumbore on the pointer tree). Returns results[N,4] int32 (+ counters[N,3])."""
        rays = _f32(rays).reshape(-1, 8)
        res = np.zeros((len(rays), 4), dtype=np.int32)
        cnt = np.zeros((len(rays), 3), dtype=np.uint32) if counters else None
        lib().orc_bvh_trace(self._h, _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0), _p(res),
                            _p(cnt) if counters else None, C.c_int(nthreads))
        return (res, cnt) if counters else res

    def compact(self):
        """CudaBVH::createCompact without the fork's random shuffle -> (nodes, woop, triIndex) int32 arrays."""
        sizes = np.zeros(3, dtype=np.int64)
        lib().orc_bvh_compact_sizes(self._h, _p(sizes))
        nodes = np.zeros(sizes[0] // 4, dtype=np.int32)
        woop = np.zeros(sizes[1] // 4, dtype=np.int32)
        idx = np.zeros(sizes[2] // 4, dtype=np.int32)
        lib().orc_bvh_compact_copy(self._h, _p(nodes), _p(woop), _p(idx))
        return nodes, woop, idx


    def basic(self, layout: int):
        """CudaBVH(bvh, layout) for layout 0..3 (AOS_AOS, AOS_SOA, SOA_AOS, SOA_SOA; CudaBVH.cpp:453-575)
        -> (nodes, woop, triIndex) int32 arrays, 4096-byte padded like the reference's, padding zero."""
        if layout not in (0, 1, 2, 3):
            raise ValueError("basic layouts are 0..3")
        sizes = np.zeros(3, dtype=np.int64)
        lib().orc_bvh_basic(self._h, C.c_int(layout), _p(sizes), None, None, None)
        nodes = np.zeros(sizes[0] // 4, dtype=np.int32)
        woop = np.zeros(sizes[1] // 4, dtype=np.int32)
        idx = np.zeros(sizes[2] // 4, dtype=np.int32)
        lib().orc_bvh_basic(self._h, C.c_int(layout), _p(sizes), _p(nodes), _p(woop), _p(idx))
        return nodes, woop, idx


def compact_trace(nodes, woop, tri_index, rays, need_closest=True, counters=False, nthreads=0):
    """CudaBVH::trace<BVHLayout_Compact> (Woop test on the flat buffers)."""
    nodes, woop, tri_index = _i32(nodes), _i32(woop), _i32(tri_index)
    rays = _f32(rays).reshape(-1, 8)
    res = np.zeros((len(rays), 4), dtype=np.int32)
    cnt = np.zeros((len(rays), 3), dtype=np.uint32) if counters else None
    lib().orc_compact_trace(_p(nodes), _p(woop), _p(tri_index), _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0),
                            _p(res), _p(cnt) if counters else None, C.c_int(nthreads))
    return (res, cnt) if counters else res


def wide4_trace(wnodes, woop, tri_index, rays, need_closest=True, counters=False, nthreads=0):
    """CPU emulation of the product's Wide4 traversal (csrc/nt_wide.cu) on its own node array; leaves / triangle test as compact_trace."""
    wnodes = np.ascontiguousarray(wnodes, dtype=np.uint32)
    woop, tri_index = _i32(woop), _i32(tri_index)
    rays = _f32(rays).reshape(-1, 8)
    res = np.zeros((len(rays), 4), dtype=np.int32)
    cnt = np.zeros((len(rays), 3), dtype=np.uint32) if counters else None
    lib().orc_wide4_trace(_p(wnodes), _p(woop), _p(tri_index), _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0),
                          _p(res), _p(cnt) if counters else None, C.c_int(nthreads))
    return (res, cnt) if counters else res


def wide4_check(wnodes, nodes, layout=4):
    """Structural check of a Wide4 array against its source Compact tree -> dict (raises on inconsistency)."""
    wnodes = np.ascontiguousarray(wnodes, dtype=np.uint32)
    nodes = _i32(nodes)
    out = np.zeros(4, dtype=np.float64)
    lib().orc_wide4_check.restype = C.c_int
    rc = lib().orc_wide4_check(_p(wnodes), C.c_int64(len(wnodes) // 16), _p(nodes), C.c_int64(nodes.nbytes), C.c_int(layout), _p(out))
    if rc != 0:
        raise AssertionError(f"Wide4 array inconsistent with its source tree (code {rc})")
    return dict(num_wide=int(out[0]), num_leaves=int(out[1]), max_depth=int(out[2]), worst_overhang_steps=float(out[3]))


def brute_trace(verts, tris, rays, need_closest=True, nthreads=0):
    verts, tris = _f32(verts).reshape(-1, 3), _i32(tris).reshape(-1, 3)
    rays = _f32(rays).reshape(-1, 8)
    res = np.zeros((len(rays), 4), dtype=np.int32)
    lib().orc_brute_trace(_p(verts), C.c_int(len(verts)), _p(tris), C.c_int(len(tris)), _p(rays), C.c_int(len(rays)),
                          C.c_int(1 if need_closest else 0), _p(res), C.c_int(nthreads))
    return res


def compact_sah(nodes, woop):
    """SAH (BVHNode.cpp:79-94 formula) of a flat Compact tree -> dict."""
    nodes, woop = _i32(nodes), _i32(woop)
    out = np.zeros(5, dtype=np.float64)
    lib().orc_compact_sah(_p(nodes), _p(woop), _p(out))
    return dict(sah=float(out[0]), num_inner=int(out[1]), num_leaf=int(out[2]), num_tris=int(out[3]), max_depth=int(out[4]))


def woopify(v0, v1, v2, gpu_form=False) -> np.ndarray:
    v = _f32(np.concatenate([v0, v1, v2]))
    out = np.zeros(12, dtype=np.float32)
    lib().orc_woopify(_p(v), _p(out), C.c_int(1 if gpu_form else 0))
    return out


def ray_triangle(v0, v1, v2, ray):
    v = _f32(np.concatenate([v0, v1, v2])); r = _f32(ray); out = np.zeros(3, dtype=np.float32)
    lib().orc_ray_triangle(_p(v), _p(r), _p(out))
    return out


def ray_triangle_woop(woop12, ray):
    w = _f32(woop12); r = _f32(ray); out = np.zeros(3, dtype=np.float32)
    lib().orc_ray_triangle_woop(_p(w), _p(r), _p(out))
    return out


def ray_box(lo, hi, ray):
    b = _f32(np.concatenate([lo, hi])); r = _f32(ray); out = np.zeros(2, dtype=np.float32)
    lib().orc_ray_box(_p(b), _p(r), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# LBVH / HLBVH
# ----------------------------------------------------------------------------------------------
def morton(verts, tris, lo, hi) -> np.ndarray:
    verts, tris = _f32(verts).reshape(-1, 3), _i32(tris).reshape(-1, 3)
    lo, hi = _f32(lo), _f32(hi)
    codes = np.zeros(len(tris), dtype=np.uint32)
    lib().orc_morton(_p(verts), C.c_int(len(verts)), _p(tris), C.c_int(len(tris)), _p(lo), _p(hi), _p(codes))
    return codes


def sort_pairs(keys, idx):
    keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    idx = _i32(idx).copy()
    lib().orc_sort_pairs(_p(keys), _p(idx), C.c_int(len(keys)))
    return keys, idx


@dataclass
class LBVH:
    nodes: np.ndarray
    woop: np.ndarray
    tri_index: np.ndarray
    sorted_keys: np.ndarray
    sorted_idx: np.ndarray
    num_nodes: int
    num_leaves: int
    num_clusters: int
    num_levels: int
    build_seconds: float


def lbvh_build(verts, tris, lo, hi, hlbvh=False, hlbvh_bits=4, leaf_size=8, epsilon=0.001) -> LBVH:
    """Restated HLBVHBuilder (serial schedule). hlbvh=False or hlbvh_bits=10 -> plain LBVH."""
    verts, tris = _f32(verts).reshape(-1, 3), _i32(tris).reshape(-1, 3)
    lo, hi = _f32(lo), _f32(hi)
    h = C.c_void_p(lib().orc_lbvh_build(_p(verts), C.c_int(len(verts)), _p(tris), C.c_int(len(tris)), _p(lo), _p(hi),
                                        C.c_int(1 if hlbvh else 0), C.c_int(hlbvh_bits), C.c_int(leaf_size), C.c_float(epsilon)))
    try:
        info = np.zeros(7, dtype=np.int64)
        sec = C.c_double(0.0)
        lib().orc_lbvh_info(h, _p(info), C.byref(sec))
        nodes = np.zeros(info[0] // 4, dtype=np.int32)
        woop = np.zeros(info[1] // 4, dtype=np.int32)
        idx = np.zeros(info[2] // 4, dtype=np.int32)
        sk = np.zeros(len(tris), dtype=np.uint32)
        si = np.zeros(len(tris), dtype=np.int32)
        lib().orc_lbvh_copy(h, _p(nodes), _p(woop), _p(idx), _p(sk), _p(si))
        return LBVH(nodes, woop, idx, sk, si, int(info[3]), int(info[4]), int(info[5]), int(info[6]), float(sec.value))
    finally:
        lib().orc_lbvh_free(h)


@dataclass
class Canonical:
    inner: np.ndarray       # [numInner, 3] = (leftTris, rightTris, word14), preorder
    boxes: np.ndarray       # [numInner, 12] node words 0..11 as floats
    leaf_sizes: np.ndarray  # [numLeaves], traversal order
    tris: np.ndarray        # concatenated leaf triangle ids, traversal order
    woop: np.ndarray        # [numTris, 12] Woop rows of the leaf triangles, traversal order


def canonical(nodes, woop, tri_index) -> Canonical:
    """Numbering-independent serialisation of a Compact tree."""
    nodes, woop, tri_index = _i32(nodes), _i32(woop), _i32(tri_index)
    sizes = np.zeros(3, dtype=np.int64)
    lib().orc_canonical(_p(nodes), _p(woop), _p(tri_index), _p(sizes), None, None, None, None, None)
    inner = np.zeros((sizes[0], 3), dtype=np.int32)
    boxes = np.zeros((sizes[0], 12), dtype=np.float32)
    ls = np.zeros(sizes[1], dtype=np.int32)
    tr = np.zeros(sizes[2], dtype=np.int32)
    wo = np.zeros((sizes[2], 12), dtype=np.float32)
    lib().orc_canonical(_p(nodes), _p(woop), _p(tri_index), _p(sizes), _p(inner), _p(boxes), _p(ls), _p(tr), _p(wo))
    return Canonical(inner, boxes, ls, tr, wo)


# ----------------------------------------------------------------------------------------------
# Ray generation
# ----------------------------------------------------------------------------------------------
def pixel_table(w, h):
    i2p = np.zeros(w * h, dtype=np.int32)
    p2i = np.zeros(w * h, dtype=np.int32)
    lib().orc_pixel_table(C.c_int(w), C.c_int(h), _p(i2p), _p(p2i))
    return i2p, p2i


def raygen_primary(origin, nscreen_to_world, w, h, max_dist, seed=0):
    origin = _f32(origin); m = _f32(nscreen_to_world).reshape(4, 4)
    rays = np.zeros((w * h, 8), dtype=np.float32)
    id2slot = np.zeros(w * h, dtype=np.int32)
    slot2id = np.zeros(w * h, dtype=np.int32)
    lib().orc_raygen_primary(_p(rays), _p(id2slot), _p(slot2id), _p(origin), _p(m), C.c_int(w), C.c_int(h), C.c_float(max_dist), C.c_uint32(seed))
    return rays, id2slot, slot2id


def raygen_ao(in_rays, in_results, normals, first, count, samples, max_dist, seed):
    in_rays = _f32(in_rays).reshape(-1, 8); in_results = _i32(in_results).reshape(-1, 4); normals = _f32(normals).reshape(-1, 3)
    out = np.zeros((count * samples, 8), dtype=np.float32)
    a = np.zeros(count * samples, dtype=np.int32); b = np.zeros(count * samples, dtype=np.int32)
    lib().orc_raygen_ao(_p(out), _p(a), _p(b), _p(in_rays), _p(in_results), _p(normals), C.c_int(first), C.c_int(count),
                        C.c_int(samples), C.c_float(max_dist), C.c_uint32(seed))
    return out, a, b


def raygen_shadow(in_rays, in_results, first, count, samples, light_pos, light_radius, seed):
    """rayGenShadowKernel (RayGenKernels.cu:240-302) -> (rays, idToSlot, slotToID)."""
    in_rays = _f32(in_rays).reshape(-1, 8); in_results = _i32(in_results).reshape(-1, 4); lp = _f32(light_pos)
    out = np.zeros((count * samples, 8), dtype=np.float32)
    a = np.zeros(count * samples, dtype=np.int32); b = np.zeros(count * samples, dtype=np.int32)
    lib().orc_raygen_shadow(_p(out), _p(a), _p(b), _p(in_rays), _p(in_results), C.c_int(first), C.c_int(count), C.c_int(samples),
                            _p(lp), C.c_float(light_radius), C.c_uint32(seed))
    return out, a, b


def ray_morton_keys(rays):
    """findAABBKernel + genMortonKeysKernel restated (RayBufferKernels.cu:70-175) -> (uint32 [N,6] keys, lo[3], hi[3])."""
    rays = _f32(rays).reshape(-1, 8)
    keys = np.zeros((len(rays), 6), dtype=np.uint32)
    aabb = np.zeros(6, dtype=np.float32)
    lib().orc_ray_morton_keys(_p(rays), C.c_int(len(rays)), _p(keys), _p(aabb))
    return keys, aabb[:3].copy(), aabb[3:].copy()


def ray_morton_order(rays, truncated=False):
    """RayBuffer::mortonSort order: order[new] = old slot (+ the 64-bit truncated keys, old-slot indexed)."""
    rays = _f32(rays).reshape(-1, 8)
    order = np.zeros(len(rays), dtype=np.int32)
    k64 = np.zeros(len(rays), dtype=np.uint64)
    lib().orc_ray_morton_order(_p(rays), C.c_int(len(rays)), C.c_int(1 if truncated else 0), _p(order), _p(k64))
    return order, k64


def count_hits(results) -> int:
    results = _i32(results).reshape(-1, 4)
    return int(lib().orc_count_hits(_p(results), C.c_int(len(results))))


def tri_normals(verts, tris) -> np.ndarray:
    verts, tris = _f32(verts).reshape(-1, 3), _i32(tris).reshape(-1, 3)
    out = np.zeros((len(tris), 3), dtype=np.float32)
    lib().orc_tri_normals(_p(verts), C.c_int(len(verts)), _p(tris), C.c_int(len(tris)), _p(out))
    return out


def invert4(m) -> np.ndarray:
    m = _f32(m).reshape(4, 4); out = np.zeros((4, 4), dtype=np.float32)
    lib().orc_invert4(_p(m), _p(out))
    return out
