"""Freezes outputs of the REFERENCE ITSELF (oracle/_ref/libref.so = the reference's own CPU sources compiled unmodified,
see oracle/Makefile) into tests/golden/ref_*.npz + ref_golden.json, so that the oracle stays pinned against the
reference where /root/reference does not exist (the GPU box, CI).

Run in the build container:  make -C oracle ref && python scripts/make_ref_golden.py
Nothing here calls the restated oracle except `canonical()` (a pure re-serialisation of the reference's node buffer,
needed because this fork shuffles Compact node order randomly, CudaBVH.cpp:65-78).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from ntrace_b200 import camera, scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def fitted_camera(verts):
    lo, hi = scenes.bbox(verts)
    c = (lo + hi) * np.float32(0.5)
    d = float(np.linalg.norm(hi - lo))
    pos = c + np.array([0.45, 0.35, 0.3], np.float32) * np.float32(d)
    return camera.look_at(pos, c, up=(0.0, 1.0, 0.0), fov=60.0, near=d * 1e-3, far=d * 4.0)


def pin_rays(verts, seed=7, n=4096):
    """128x96 primary rays (oracle raygen is input data here, stored by sha) + seeded incoherent rays with finite tmax."""
    cam = fitted_camera(verts)
    prim, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    lo, hi = scenes.bbox(verts)
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    tmax = (rng.uniform(0.05, 1.5, (n, 1)) * np.linalg.norm(hi - lo)).astype(np.float32)
    rnd = np.concatenate([o, np.zeros((n, 1), np.float32), d, tmax], axis=1).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([prim, rnd], axis=0))


CONFIGS = {  # name -> (split, minLeaf, maxLeaf)
    "sah_1_1": (False, 1, 1), "sah_1_8": (False, 1, 8), "split_1_1": (True, 1, 1), "split_1_8": (True, 1, 8),
}


def main():
    assert ref.build() is not None, "reference tree not present"
    os.makedirs(OUT, exist_ok=True)
    meta = {}
    for name in ("map", "head", "room5000_seed21"):
        if name.startswith("room"):
            verts, tris = scenes.room(5_000, seed=21)
        else:
            d = np.load(os.path.join(OUT, f"{name}.npz"))
            verts, tris = d["verts"], d["tris"]
        rays = pin_rays(verts)
        arrays = {}
        meta[name] = {"rays_sha": sha(rays), "num_rays": int(len(rays)), "configs": {}}
        for cfg, (split, mn, mx) in CONFIGS.items():
            b = ref.RefBVH(verts, tris, split=split, min_leaf=mn, max_leaf=mx, split_alpha=1.0e-5)
            st = b.stats()
            nodes, woop, idx = b.compact()
            c = oracle.canonical(nodes, woop, idx)
            tree = b.trace(rays, True)
            flat = b.compact_trace(rays, True)
            flat_any = b.compact_trace(rays, False)
            tree_any = b.trace(rays, False)
            aos = b.layout_trace(0, rays, True)
            assert np.array_equal(aos[:, :2], flat[:, :2]), "CudaBVH::trace<AOS_AOS> and <Compact> disagree"
            layout_sha = {str(L): [sha(a) for a in b.layout(L)] for L in range(4)}
            meta[name]["configs"][cfg] = dict(st, layout_sha=layout_sha, aos_aos_trace_equals_compact=True, tri_indices_sha=sha(b.tri_indices()), woop_buffer_sha=sha(woop), tri_index_buffer_sha=sha(idx),
                                              inner_sha=sha(c.inner), boxes_sha=sha(c.boxes), leaf_sizes_sha=sha(c.leaf_sizes),
                                              hits=int((flat[:, 0] >= 0).sum()))
            arrays[f"{cfg}.tree"] = tree[:, :2].copy()
            arrays[f"{cfg}.flat"] = flat[:, :2].copy()
            arrays[f"{cfg}.tree_any"] = tree_any[:, :2].copy()
            arrays[f"{cfg}.flat_any"] = flat_any[:, :2].copy()
            print(name, cfg, st)
        np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **arrays)
    # primitive known-answers: random boxes / triangles / rays through Intersect::* and invert()
    rng = np.random.default_rng(123)
    n = 2000
    prim_rays = np.concatenate([rng.normal(size=(n, 3)), np.zeros((n, 1)), rng.normal(size=(n, 3)), rng.uniform(1, 50, (n, 1))], 1).astype(np.float32)
    blo = rng.normal(size=(n, 3)).astype(np.float32); bhi = blo + rng.uniform(0, 3, (n, 3)).astype(np.float32)
    tv = rng.normal(size=(n, 9)).astype(np.float32) * np.float32(2.0)
    mats = rng.normal(size=(64, 4, 4)).astype(np.float32)
    box_out = np.stack([ref.ray_box(blo[i], bhi[i], prim_rays[i]) for i in range(n)])
    tri_out = np.stack([ref.ray_triangle(tv[i, 0:3], tv[i, 3:6], tv[i, 6:9], prim_rays[i]) for i in range(n)])
    inv_out = np.stack([ref.invert4(m) for m in mats])
    np.savez_compressed(os.path.join(OUT, "ref_primitives.npz"), rays=prim_rays, blo=blo, bhi=bhi, tv=tv, mats=mats,
                        box_out=box_out, tri_out=tri_out, inv_out=inv_out)
    meta["pixel_table_sha"] = {f"{w}x{h}": [sha(a) for a in ref.pixel_table(w, h)] for w, h in ((1024, 768), (100, 75), (37, 21), (8, 8), (7, 5))}
    json.dump(meta, open(os.path.join(OUT, "ref_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
