"""Generates tests/golden/: the two real meshes shipped with the reference (converted with the reference's OBJ
ingestion rule) and known-answer values computed by the CPU oracle on them.

Run in the build container (needs /root/reference): python scripts/make_golden.py
The reference ships no golden vectors for this path, so these pin the *oracle* against drift and
give the GPU tests fixed inputs; the pin against the reference itself is scripts/make_ref_golden.py + tests/test_reference_pin.py.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ntrace_b200 import camera, mesh_io, scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def fitted_camera(verts):
    lo, hi = scenes.bbox(verts)
    c = (lo + hi) * np.float32(0.5)
    d = float(np.linalg.norm(hi - lo))
    pos = c + np.array([0.45, 0.35, 0.3], np.float32) * np.float32(d)
    return camera.look_at(pos, c, up=(0.0, 1.0, 0.0), fov=60.0, near=d * 1e-3, far=d * 4.0)


def golden_for(verts, tris):
    lo, hi = scenes.bbox(verts)
    g = {"num_verts": int(len(verts)), "num_tris": int(len(tris)), "bbox": [lo.tolist(), hi.tolist()]}
    codes = oracle.morton(verts, tris, lo, hi)
    g["morton_sha"] = sha(codes)
    for leaf in (8, 1):
        r = oracle.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf)
        c = oracle.canonical(r.nodes, r.woop, r.tri_index)
        s = oracle.compact_sah(r.nodes, r.woop)
        g[f"lbvh_leaf{leaf}"] = {"sorted_idx_sha": sha(r.sorted_idx), "tree_sha": sha(c.inner), "leaf_sizes_sha": sha(c.leaf_sizes),
                                 "boxes_sha": sha(c.boxes), "woop_sha": sha(c.woop), "num_nodes": r.num_nodes, "num_leaves": r.num_leaves,
                                 "sah": s["sah"], "max_depth": s["max_depth"]}
    r = oracle.lbvh_build(verts, tris, lo, hi, hlbvh=True, hlbvh_bits=4, leaf_size=8)
    s = oracle.compact_sah(r.nodes, r.woop)
    g["hlbvh_bits4_leaf8"] = {"num_clusters": r.num_clusters, "num_nodes": r.num_nodes, "num_leaves": r.num_leaves, "sah": s["sah"]}
    cam = fitted_camera(verts)
    w, h = 128, 96
    rays, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    g["rays_sha"] = sha(rays)
    for name, b in (("sahbvh", oracle.BUILDER_SAH), ("splitbvh", oracle.BUILDER_SPLIT)):
        bvh = oracle.CpuBVH(verts, tris, b, 1, 1, 1.0e-5)
        st = bvh.stats()
        res = bvh.trace(rays, True)
        nodes, woop, idx = bvh.compact()
        flat = oracle.compact_trace(nodes, woop, idx, rays, True)
        hit = res[:, 0] >= 0
        g[name] = {"sah": st.sah, "num_inner": st.num_inner, "num_leaf": st.num_leaf, "duplicates": st.duplicates, "max_depth": st.max_depth,
                   "nodes_sha": sha(nodes), "tree_trace_ids_sha": sha(res[:, 0]), "flat_trace_ids_sha": sha(flat[:, 0]),
                   "hits": int(hit.sum()), "sum_t": float(res[hit, 1].view(np.float32).astype(np.float64).sum())}
    return g


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {}
    for name, rel in (("map", "data/models/Map/Map.obj"), ("head", "data/models/Head/head.obj")):
        v, t = mesh_io.load_obj(os.path.join("/root/reference", rel))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), verts=v, tris=t)
        gold[name] = golden_for(v, t)
        print(name, gold[name]["num_tris"], gold[name]["splitbvh"]["sah"])
    v, t = scenes.room(5_000, seed=21)
    gold["room5000_seed21"] = golden_for(v, t)
    gold["pixel_table_sha"] = {f"{w}x{h}": sha(oracle.pixel_table(w, h)[0]) for w, h in ((1024, 768), (100, 75), (37, 21))}
    json.dump(gold, open(os.path.join(OUT, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
