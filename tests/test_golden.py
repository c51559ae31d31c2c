"""Known-answer tests on committed fixtures (tests/golden/, made by scripts/make_golden.py): the two real meshes
the reference ships (Head 18,678 tris, Map 488 tris) and a seeded synthetic room.  They pin the oracle against
drift; the GPU variants check the CUDA path against the same committed answers."""
import hashlib
import json
import os

import numpy as np
import pytest

from ntrace_b200 import camera, scenes

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(HERE, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def load(name):
    if name.startswith("room"):
        return scenes.room(5_000, seed=21)
    d = np.load(os.path.join(HERE, f"{name}.npz"))
    return d["verts"], d["tris"]


def fitted_camera(verts):
    lo, hi = scenes.bbox(verts)
    c = (lo + hi) * np.float32(0.5)
    d = float(np.linalg.norm(hi - lo))
    pos = c + np.array([0.45, 0.35, 0.3], np.float32) * np.float32(d)
    return camera.look_at(pos, c, up=(0.0, 1.0, 0.0), fov=60.0, near=d * 1e-3, far=d * 4.0)


NAMES = ["map", "head", "room5000_seed21"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(orc, name):
    verts, tris = load(name)
    g = GOLD[name]
    assert len(tris) == g["num_tris"] and len(verts) == g["num_verts"]
    lo, hi = scenes.bbox(verts)
    assert sha(orc.morton(verts, tris, lo, hi)) == g["morton_sha"]
    for leaf in (8, 1):
        r = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf)
        c = orc.canonical(r.nodes, r.woop, r.tri_index)
        e = g[f"lbvh_leaf{leaf}"]
        assert (sha(r.sorted_idx), sha(c.inner), sha(c.leaf_sizes), sha(c.boxes), sha(c.woop)) == \
               (e["sorted_idx_sha"], e["tree_sha"], e["leaf_sizes_sha"], e["boxes_sha"], e["woop_sha"])
        assert (r.num_nodes, r.num_leaves) == (e["num_nodes"], e["num_leaves"])
    cam = fitted_camera(verts)
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    assert sha(rays) == g["rays_sha"]
    for key, b in (("sahbvh", orc.BUILDER_SAH), ("splitbvh", orc.BUILDER_SPLIT)):
        if name == "head" and key == "splitbvh":
            continue                       # ~10 s; covered by map/room and by the gpu suite
        bvh = orc.CpuBVH(verts, tris, b, 1, 1, 1.0e-5)
        st = bvh.stats()
        e = g[key]
        assert (st.num_inner, st.num_leaf, st.duplicates, st.max_depth) == (e["num_inner"], e["num_leaf"], e["duplicates"], e["max_depth"])
        assert abs(st.sah - e["sah"]) <= 1e-5 * e["sah"]
        res = bvh.trace(rays, True)
        assert sha(res[:, 0]) == e["tree_trace_ids_sha"] and int((res[:, 0] >= 0).sum()) == e["hits"]


def test_pixel_table_golden(orc):
    for k, v in GOLD["pixel_table_sha"].items():
        w, h = map(int, k.split("x"))
        assert sha(orc.pixel_table(w, h)[0]) == v


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(gpu_host, orc, name):
    from ntrace_b200 import capi
    verts, tris = load(name)
    g = GOLD[name]
    lo, hi = scenes.bbox(verts)
    for leaf in (8, 1):
        capi.bvh_build(capi.BUILDER_LBVH, np.ascontiguousarray(verts), np.ascontiguousarray(tris), lo, hi, 10, leaf, 0.001)
        nodes, woop, idx, _ = capi.bvh_download()
        keys, order = capi.bvh_build_debug(len(tris))
        c = orc.canonical(nodes, woop, idx)
        e = g[f"lbvh_leaf{leaf}"]
        assert sha(order) == e["sorted_idx_sha"] and sha(c.inner) == e["tree_sha"] and sha(c.leaf_sizes) == e["leaf_sizes_sha"]
        assert sha(c.woop) == e["woop_sha"]
        assert np.array_equal(c.boxes, orc.canonical(*_ref_lbvh(orc, verts, tris, lo, hi, leaf)).boxes)
        assert abs(orc.compact_sah(nodes, woop)["sah"] - e["sah"]) <= 0.005 * e["sah"]
    capi.bvh_build(capi.BUILDER_HLBVH, np.ascontiguousarray(verts), np.ascontiguousarray(tris), lo, hi, 4, 8, 0.001)
    nodes, woop, idx, _ = capi.bvh_download()
    e = g["hlbvh_bits4_leaf8"]
    assert len(nodes) // 16 == e["num_nodes"] and abs(orc.compact_sah(nodes, woop)["sah"] - e["sah"]) <= 0.005 * e["sah"]
    # trace the reference-built SAH BVH with the CUDA kernel on the golden rays
    cam = fitted_camera(verts)
    rays = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    assert sha(rays.rays_host()) == g["rays_sha"]
    bvh = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1, 1.0e-5)
    nodes, woop, idx = bvh.compact()
    assert sha(nodes) == g["sahbvh"]["nodes_sha"]
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(gpu_host.CudaBVH(nodes, woop, idx))
    tracer.traceBatch(rays)
    assert sha(rays.results_host()[:, 0]) == g["sahbvh"]["flat_trace_ids_sha"]


def _ref_lbvh(orc, verts, tris, lo, hi, leaf):
    r = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf)
    return r.nodes, r.woop, r.tri_index
