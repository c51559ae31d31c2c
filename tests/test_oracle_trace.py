"""The oracle cross-validated three ways: pointer-tree BVH::trace (Moller-Trumbore) vs flat
CudaBVH::trace (Woop) vs brute force over all triangles, plus unit checks of the intersection primitives."""
import numpy as np
import pytest

from ntrace_b200 import camera, scenes


@pytest.fixture(scope="module")
def scene(orc):
    verts, tris = scenes.room(8_000, seed=5, wall_frac=0.3)
    cam = camera.named_camera("conference")
    w, h = 160, 120
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    return verts, tris, rays


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


@pytest.mark.parametrize("builder", ["sah", "split"])
def test_tree_trace_equals_brute_force(orc, scene, builder):
    verts, tris, rays = scene
    bvh = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if builder == "split" else orc.BUILDER_SAH, 1, 1)
    got = bvh.trace(rays, True)
    ref = orc.brute_trace(verts, tris, rays, True)
    same = got[:, 0] == ref[:, 0]
    assert same.mean() >= 0.9999
    # same arithmetic (Moller-Trumbore on the same vertices) -> t is bit-identical where ids agree
    assert np.array_equal(got[same, 1], ref[same, 1])
    # mismatches are exact ties in t (first-tested-wins)
    assert np.array_equal(got[~same, 1], ref[~same, 1])


def test_flat_woop_trace_matches_tree_trace(orc, scene):
    verts, tris, rays = scene
    bvh = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT, 1, 1)
    nodes, woop, idx = bvh.compact()
    flat, cnt = orc.compact_trace(nodes, woop, idx, rays, True, counters=True)
    tree, cnt2 = bvh.trace(rays, True, counters=True)
    same = flat[:, 0] == tree[:, 0]
    assert same.mean() >= 0.999
    hit = same & (tree[:, 0] >= 0)
    assert _rel(flat[hit, 1].view(np.float32), tree[hit, 1].view(np.float32)).max() <= 1e-5
    # both walk the same tree with the same near/far rule: node / triangle test counts agree almost everywhere
    assert (cnt[:, 0] == cnt2[:, 0]).mean() >= 0.99
    # miss keeps t = tmax in the flat tracer (CudaBVH.cpp:271-272)
    miss = flat[:, 0] < 0
    assert np.array_equal(flat[miss, 1].view(np.float32), rays[miss, 7])


def test_anyhit_agrees_on_hit_or_miss(orc, scene):
    verts, tris, rays = scene
    bvh = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT, 1, 1)
    nodes, woop, idx = bvh.compact()
    short = rays.copy()
    short[:, 7] = 12.0                                     # clip rays so that some miss
    closest = orc.compact_trace(nodes, woop, idx, short, True)
    anyhit = orc.compact_trace(nodes, woop, idx, short, False)
    assert np.array_equal(closest[:, 0] >= 0, anyhit[:, 0] >= 0)
    assert 0.05 < (closest[:, 0] >= 0).mean() < 0.95
    tree_any = bvh.trace(short, False)
    assert ((tree_any[:, 0] >= 0) == (anyhit[:, 0] >= 0)).mean() >= 0.9999


def test_ray_primitives(orc):
    v0, v1, v2 = np.array([0, 0, 0], np.float32), np.array([1, 0, 0], np.float32), np.array([0, 1, 0], np.float32)
    ray = np.array([0.25, 0.25, 1, 0, 0, 0, -1, 10], np.float32)
    t, u, v = orc.ray_triangle(v0, v1, v2, ray)
    assert t == 1.0 and u == 0.25 and v == 0.25
    for gpu_form in (False, True):
        w = orc.woopify(v0, v1, v2, gpu_form)
        tw, uw, vw = orc.ray_triangle_woop(w, ray)
        assert abs(tw - 1.0) < 1e-6 and abs(uw - 0.5) < 1e-6 and abs(vw - 0.25) < 1e-6  # Woop (u,v) weight v0,v1; MT weights v1,v2
    # strict interval: t == tmax is rejected (Util.cpp:86,109)
    ray2 = ray.copy(); ray2[7] = 1.0
    assert orc.ray_triangle(v0, v1, v2, ray2)[0] > 1e30
    assert orc.ray_triangle_woop(orc.woopify(v0, v1, v2), ray2)[0] > 1e30
    # back face is accepted (two-sided, EPSILON = 0)
    ray3 = np.array([0.25, 0.25, -1, 0, 0, 0, 1, 10], np.float32)
    assert orc.ray_triangle(v0, v1, v2, ray3)[0] == 1.0
    # slab test
    tmin, tmax = orc.ray_box([0, 0, 0], [1, 1, 1], np.array([-1, 0.5, 0.5, 0, 1, 0, 0, 10], np.float32))
    assert tmin == 1.0 and tmax == 2.0
    # the two Woop constructions agree closely (different formulas: 4x4 cofactor inverse vs 3x3 adjugate)
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b, c = rng.normal(size=(3, 3)).astype(np.float32)
        w1, w2 = orc.woopify(a, b, c, False), orc.woopify(a, b, c, True)
        assert np.allclose(w1, w2, rtol=2e-3, atol=2e-4)


def test_woop_terminator_never_aliased(orc, scene):
    verts, tris, _ = scene
    nodes, woop, idx = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1).compact()
    w = woop.reshape(-1, 4)
    NEG0 = np.int32(-2147483648)
    # walk the array the way the kernels do: triangles are 3 float4 apart, a leaf ends at a float4 whose .x is -0
    pos, n_term, n_tri = 0, 0, 0
    while pos < len(w):
        if w[pos, 0] == NEG0:
            assert (w[pos] == NEG0).all() and idx[pos] == 0       # all four lanes set (Math.hpp:281), index 0
            n_term += 1
            pos += 1
        else:
            n_tri += 1                                            # row 0 of a triangle never has .x == -0 (CudaBVH.cpp:627-628)
            pos += 3
    assert n_tri == len(tris) and n_term == len(tris)             # leaf preference (1,1): one terminator per triangle
    assert len(idx) == len(w)
