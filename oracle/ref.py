"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/_ref/libref.so``: the reference's OWN CPU sources for the path
(SAHBVHBuilder, SplitBVHBuilder, BVHNode SAH metric, BVH::trace, CudaBVH::createCompact /
woopifyTri / trace, Intersect::*), compiled unmodified from ``/root/reference`` by
``make -C oracle ref`` (see ``oracle/Makefile`` and ``oracle/ref_shim/``).

Used to pin the restatement (``oracle/liborc.so``): ``tests/test_reference_pin.py`` compares the two
live, ``scripts/make_golden.py`` freezes reference outputs into ``tests/golden/ref_*.npz`` so the
pin also holds where ``/root/reference`` is absent, and ``bench.py --impl reference`` times it.
Nothing under ``ntrace_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_ROOT = os.environ.get("NTRACE_REFERENCE_ROOT", "/root/reference")
_lib = None


def build(force: bool = False) -> str | None:
    """Build ``_ref/libref.so`` when the reference tree is present; return its path or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "rt")):
        shim = os.path.join(_HERE, "ref_shim")
        srcs = [os.path.join(_HERE, "Makefile")] + [os.path.join(d, f) for d, _, fs in os.walk(shim) for f in fs]
        stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
        if force or stale:
            subprocess.run(["make", "-C", _HERE, "-B", "ref", f"REF={REFERENCE_ROOT}"], check=True, capture_output=True)
    return LIB_PATH if os.path.exists(LIB_PATH) else None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/libref.so not built (needs /root/reference; run `make -C oracle ref`)")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_build.restype = C.c_void_p
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class RefBVH:
    """``FW::BVH`` built by the reference's SAHBVHBuilder (split=False) or SplitBVHBuilder (split=True)."""

    def __init__(self, verts, tris, split=True, min_leaf=1, max_leaf=1, split_alpha=1.0e-5):
        self.verts = _f32(verts).reshape(-1, 3)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        self._h = C.c_void_p(lib().ref_build(_p(self.verts), C.c_int(len(self.verts)), _p(self.tris), C.c_int(len(self.tris)),
                                             C.c_int(1 if split else 0), C.c_int(min_leaf), C.c_int(max_leaf), C.c_float(split_alpha)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_free(self._h)
            self._h = None

    def stats(self) -> dict:
        """BVH::Stats as BVH::BVH fills it (BVH.cpp:73-82)."""
        out = np.zeros(6, dtype=np.float64)
        lib().ref_stats(self._h, _p(out))
        return dict(sah=float(out[0]), num_inner=int(out[1]), num_leaf=int(out[2]), num_tris=int(out[3]), max_depth=int(out[4]),
                    num_tri_indices=int(out[5]))

    def tri_indices(self) -> np.ndarray:
        out = np.zeros(self.stats()["num_tri_indices"], dtype=np.int32)
        lib().ref_tri_indices(self._h, _p(out))
        return out

    def trace(self, rays, need_closest=True) -> np.ndarray:
        """BVH::trace (BVH.cpp:90-110)."""
        rays = _f32(rays).reshape(-1, 8)
        res = np.zeros((len(rays), 4), dtype=np.int32)
        lib().ref_trace(self._h, _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0), _p(res))
        return res

    def compact(self):
        """CudaBVH(bvh, BVHLayout_Compact) -> (nodes, woop, triIndex) int32 arrays; node order is the fork's random shuffle."""
        sizes = np.zeros(3, dtype=np.int64)
        lib().ref_compact_sizes(self._h, _p(sizes))
        nodes = np.zeros(sizes[0] // 4, dtype=np.int32)
        woop = np.zeros(sizes[1] // 4, dtype=np.int32)
        idx = np.zeros(sizes[2] // 4, dtype=np.int32)
        lib().ref_compact_copy(self._h, _p(nodes), _p(woop), _p(idx))
        return nodes, woop, idx

    def layout(self, layout: int):
        """CudaBVH(bvh, layout) for any BVHLayout value (0 AOS_AOS, 1 AOS_SOA, 2 SOA_AOS, 3 SOA_SOA, 4 Compact, 5 Compact2)
        -> (nodes, woop, triIndex) int32 arrays, 4096-byte padded as the reference allocates them (padding zero-filled)."""
        sizes = np.zeros(3, dtype=np.int64)
        if lib().ref_layout_sizes(self._h, C.c_int(layout), _p(sizes)):
            raise ValueError("bad layout")
        nodes = np.zeros(sizes[0] // 4, dtype=np.int32)
        woop = np.zeros(sizes[1] // 4, dtype=np.int32)
        idx = np.zeros(sizes[2] // 4, dtype=np.int32)
        lib().ref_layout_copy(self._h, C.c_int(layout), _p(nodes), _p(woop), _p(idx))
        return nodes, woop, idx

    def serialize(self, layout: int = 4) -> bytes:
        """CudaBVH(bvh, layout).serialize() through the reference's own stream operators -> the bvhcache byte stream."""
        lib().ref_serialize.restype = C.c_longlong
        n = lib().ref_serialize(self._h, C.c_int(layout), None, C.c_longlong(0))
        if n < 0:
            raise ValueError("bad layout")
        buf = np.zeros(n, dtype=np.uint8)
        lib().ref_serialize(self._h, C.c_int(layout), _p(buf), C.c_longlong(n))
        return buf.tobytes()

    def layout_trace(self, layout: int, rays, need_closest=True) -> np.ndarray:
        """CudaBVH::trace on the AOS_AOS (0) or Compact (4) buffers."""
        rays = _f32(rays).reshape(-1, 8)
        res = np.zeros((len(rays), 4), dtype=np.int32)
        if lib().ref_layout_trace(self._h, C.c_int(layout), _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0), _p(res)):
            raise ValueError("CudaBVH::trace handles AOS_AOS and Compact only")
        return res

    def compact_trace(self, rays, need_closest=True, nthreads=1) -> np.ndarray:
        """CudaBVH::trace on the Compact layout (CudaBVH.cpp:213-302)."""
        rays = _f32(rays).reshape(-1, 8)
        res = np.zeros((len(rays), 4), dtype=np.int32)
        if nthreads <= 1:
            lib().ref_compact_trace(self._h, _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0), _p(res))
        else:
            lib().ref_compact_trace_mt(self._h, _p(rays), C.c_int(len(rays)), C.c_int(1 if need_closest else 0), _p(res), C.c_int(nthreads))
        return res


def deserialize(data: bytes):
    """CudaBVH(InputStream&) of the reference on a byte stream -> (layout, nodes, woop, triIndex)."""
    raw = np.frombuffer(data, dtype=np.uint8).copy()
    sizes = np.zeros(4, dtype=np.int64)
    if lib().ref_deserialize(_p(raw), C.c_longlong(len(raw)), _p(sizes), None, None, None):
        raise ValueError("the reference's reader rejected the stream")
    nodes = np.zeros(sizes[1] // 4, dtype=np.int32)
    woop = np.zeros(sizes[2] // 4, dtype=np.int32)
    idx = np.zeros(sizes[3] // 4, dtype=np.int32)
    lib().ref_deserialize(_p(raw), C.c_longlong(len(raw)), _p(sizes), _p(nodes), _p(woop), _p(idx))
    return int(sizes[0]), nodes, woop, idx


def ray_box(lo, hi, ray):
    b = _f32(np.concatenate([lo, hi])); r = _f32(ray); out = np.zeros(2, dtype=np.float32)
    lib().ref_ray_box(_p(b), _p(r), _p(out))
    return out


def ray_triangle(v0, v1, v2, ray):
    v = _f32(np.concatenate([v0, v1, v2])); r = _f32(ray); out = np.zeros(3, dtype=np.float32)
    lib().ref_ray_triangle(_p(v), _p(r), _p(out))
    return out


def ray_triangle_woop(woop12, ray):
    w = _f32(woop12); r = _f32(ray); out = np.zeros(3, dtype=np.float32)
    lib().ref_ray_triangle_woop(_p(w), _p(r), _p(out))
    return out


def pixel_table(w, h):
    """PixelTable::setSize (PixelTable.cpp:57-141) -> (indexToPixel, pixelToIndex)."""
    i2p = np.zeros(w * h, dtype=np.int32); p2i = np.zeros(w * h, dtype=np.int32)
    lib().ref_pixel_table(C.c_int(w), C.c_int(h), _p(i2p), _p(p2i))
    return i2p, p2i


def invert4(m):
    a = _f32(m).reshape(4, 4); out = np.zeros((4, 4), dtype=np.float32)
    lib().ref_invert4(_p(a), _p(out))
    return out
