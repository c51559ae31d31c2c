"""Freezes what the REFERENCE ITSELF (oracle/_ref/libref.so, compiled unmodified: oracle/Makefile) produces for the bvhcache path:
  tests/golden/ref_map_compact.dat   CudaBVH(BVH(Map.obj, SplitBVH, leaf 1/1), BVHLayout_Compact).serialize() through the reference's own
                                     stream operators (CudaBVH.cpp:116-125, io/Stream.cpp) — the byte stream of bvhcache/*.dat files
  tests/golden/ref_cache_golden.json known answers of FW::hashBuffer / hashBits / Scene::hash / the cache-name formula (Hash.cpp, Renderer.cpp:173-178)
Run in the build container:  make -C oracle ref && python scripts/make_ref_cache_golden.py"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from ntrace_b200 import mesh_io  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    v, t = mesh_io.load_obj("/root/reference/data/models/Map/Map.obj")
    r = ref.RefBVH(v, t, split=True, min_leaf=1, max_leaf=1)
    data = r.serialize(4)
    open(os.path.join(OUT, "ref_map_compact.dat"), "wb").write(data)
    lib = ref.lib()
    for f in ("ref_hash_buffer", "ref_cache_name_hash", "ref_scene_hash"):
        getattr(lib, f).restype = C.c_uint
    buf = np.frombuffer(data, dtype=np.uint8)
    sizes = [0, 1, 2, 3, 4, 5, 11, 12, 13, 23, 24, 25, 1000, 1001, len(data)]
    kat = {"dat_sha256": hashlib.sha256(data).hexdigest(),
           "hash_buffer": {str(n): int(lib.ref_hash_buffer(C.c_void_p(buf.ctypes.data), C.c_int(n))) for n in sizes},
           "hash_buffer_unaligned": {str(n): int(lib.ref_hash_buffer(C.c_void_p(buf.ctypes.data + 1), C.c_int(n))) for n in (4, 12, 1000)},
           "scene_hash": {"args": [1, 2, 3, 4, 5], "value": int(lib.ref_scene_hash(1, 2, 3, 4, 5))},
           "cache_name_hash": [{"sceneHash": s, "minLeaf": a, "maxLeaf": b, "splitAlpha": 1.0e-5, "layout": L, "ds": ds,
                                "value": int(lib.ref_cache_name_hash(C.c_uint(s), a, b, C.c_float(1.0e-5), L, ds.encode()))}
                               for s, a, b, L, ds in ((0x12345678, 1, 1, 4, "BVH"), (0xdeadbeef, 1, 1, 5, "BVH"), (7, 1, 8, 4, "KDTree"))]}
    json.dump(kat, open(os.path.join(OUT, "ref_cache_golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(kat)[:300])


if __name__ == "__main__":
    main()
