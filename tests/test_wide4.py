"""Wide4 (4-wide, 8-bit quantised node array derived from the reference's Compact CudaBVH, csrc/nt_wide.cu), CPU side:
the host conversion (a C-ABI entry point that touches no device) checked structurally and, through the oracle's emulation of
the product's traversal, against the reference's flat Compact tracer on the same BVH and rays.

What the Wide4 path promises (include/ntrace_b200.h, nt_set_kernel): the same leaves and triangle test as the binary kernels, so
closest-hit t is bit-identical and ids differ only on exact-t ties (north_star: >= 99.99 % ids, mismatches only within 1e-4 rel t);
any-hit rays report A hit — the hit/miss flag is the reference's, the triangle may be another one (different visiting order)."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes


def _rays(orc, verts, tris, w=192, h=144):
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far, 0)
    return cam, rays


def _compare(orc, nodes, woop, idx, wn, rays, closest, scene=None):
    ref = orc.compact_trace(nodes, woop, idx, rays, closest)
    got = orc.wide4_trace(wn, woop, idx, rays, closest)
    hit_r, hit_g = ref[:, 0] >= 0, got[:, 0] >= 0
    assert (hit_r == hit_g).mean() >= 0.9999                     # boxes are conservative: the same rays hit something
    if closest:
        same = ref[:, 0] == got[:, 0]
        assert same.mean() >= 0.9999, same.mean()
        tr, tg = ref[:, 1].view(np.float32), got[:, 1].view(np.float32)
        both = hit_r & hit_g
        assert (ref[both, 1] == got[both, 1]).mean() >= 0.9999   # same triangle arithmetic: t bit-identical
        mm = ~same & both
        if mm.any():
            assert (np.abs(tr[mm] - tg[mm]) <= 1e-4 * np.abs(tr[mm])).all()
    else:
        # every reported hit must be a real hit of that triangle: the same (ray, triangle) test gives the same t bits
        sub = np.nonzero(hit_g)[0][:2000]
        # an any-hit t can never be below the closest-hit t of the same ray ...
        closest_t = orc.compact_trace(nodes, woop, idx, rays[sub], True)[:, 1].view(np.float32)
        assert (got[sub, 1].view(np.float32) >= closest_t).all()
        # ... and the reported triangle must really be hit at the reported t (Moeller-Trumbore on its raw vertices)
        if scene is not None:
            verts, tris = scene
            for i in sub[:300]:
                tri = tris[got[i, 0]]
                t, _, _ = orc.ray_triangle(verts[tri[0]], verts[tri[1]], verts[tri[2]], rays[i])
                assert abs(t - got[i, 1:2].view(np.float32)[0]) <= 1e-4 * abs(t) + 1e-6
    return ref, got


@pytest.mark.parametrize("builder_leaf", [("split", 1), ("sah", 8)])
def test_wide4_conversion_and_emulation_match_compact_tracer(orc, builder_leaf):
    builder, max_leaf = builder_leaf
    verts, tris = scenes.room(12_000, seed=5, wall_frac=0.3)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT if builder == "split" else orc.BUILDER_SAH, 1, max_leaf)
    nodes, woop, idx = cpu.compact()
    wn, depth = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes, woop.nbytes)
    info = orc.wide4_check(wn, nodes, 4)
    assert info["num_wide"] == len(wn) // 16 and info["max_depth"] == depth
    assert info["num_wide"] < 0.62 * (len(nodes) // 16)          # a wide node folds a binary node and (most of) its children
    assert info["worst_overhang_steps"] <= 1.3                   # outward rounding: at most one grid step + the 1/8-step margins
    cam, rays = _rays(orc, verts, tris)
    ref, _ = _compare(orc, nodes, woop, idx, wn, rays, True)
    normals = orc.tri_normals(verts, tris)
    for dist, closest in ((5.0, False), (cam.far, True)):
        sec, _, _ = orc.raygen_ao(rays, ref, normals, 0, 4096, 8, dist, 0x9E3779B9)
        _compare(orc, nodes, woop, idx, wn, sec, closest, scene=(verts, tris))
    # fewer node visits is the point of the wide form
    _, c2 = orc.compact_trace(nodes, woop, idx, rays, True, counters=True)
    _, c4 = orc.wide4_trace(wn, woop, idx, rays, True, counters=True)
    assert c4[:, 0].mean() < 0.65 * c2[:, 0].mean()


def test_wide4_from_compact2_gives_the_same_nodes(orc):
    verts, tris = scenes.room(6_000, seed=3)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 4)
    nodes, woop, idx = cpu.compact()
    wn4, d4 = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes, woop.nbytes)
    # Compact2 = the same tree with inner links / 16 (createCompact(bvh, 16), CudaBVH.cpp:86,614)
    n2 = nodes.copy().reshape(-1, 16)
    for c in (12, 13):
        inner = n2[:, c] >= 0
        n2[inner, c] //= 16
    wn5, d5 = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT2, n2.reshape(-1), woop.nbytes)
    assert d4 == d5 and np.array_equal(wn4, wn5)
    orc.wide4_check(wn5, n2.reshape(-1), 5)


def test_wide4_conversion_refuses_malformed_and_too_deep_trees():
    # a chain 200 levels deep: child 0 = next node, child 1 = an (empty) leaf
    depth = 200
    nodes = np.zeros((depth, 16), dtype=np.int32)
    f = nodes.view(np.float32)
    f[:, 0:12:2] = 0.0
    f[:, 1:12:2] = 1.0
    for i in range(depth):
        nodes[i, 12] = (i + 1) * 64 if i + 1 < depth else ~0
        nodes[i, 13] = ~0
    woop = np.full(4, np.int32(-2**31), dtype=np.int32)
    with pytest.raises(capi.NtError, match="too deep"):
        capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes.reshape(-1), woop.nbytes)
    bad = nodes[:4].copy()
    bad[3, 12] = 64                                       # cycle
    with pytest.raises(capi.NtError, match="cycle|outside"):
        capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, bad.reshape(-1), woop.nbytes)
    bad = nodes[:4].copy()
    bad[2, 12] = 4096                                     # link past the end
    with pytest.raises(capi.NtError, match="outside"):
        capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, bad.reshape(-1), woop.nbytes)
    bad = nodes[:4].copy()
    bad[3, 12] = ~50                                      # leaf link past the triangle buffer
    with pytest.raises(capi.NtError, match="outside"):
        capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, bad.reshape(-1), woop.nbytes)
    with pytest.raises(capi.NtError):
        capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes.reshape(-1)[:8], woop.nbytes)


def test_wide4_flat_and_degenerate_boxes_stay_conservative(orc):
    # axis-aligned quads (zero-thickness boxes, exact binary fractions and awkward offsets), plus a degenerate sliver
    v = []
    t = []
    for k, z in enumerate([0.0, 0.1, 1.0 / 3.0, 7.25, -3.3333333]):
        b = len(v)
        v += [(0 + k, 0, z), (1 + k, 0, z), (1 + k, 1, z), (0 + k, 1, z)]
        t += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    b = len(v)
    v += [(0, 0, 5), (1e-7, 0, 5), (0, 1e-7, 5)]
    t += [(b, b + 1, b + 2)]
    verts = np.array(v, dtype=np.float32)
    tris = np.array(t, dtype=np.int32)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SAH, 1, 1)
    nodes, woop, idx = cpu.compact()
    wn, _ = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes, woop.nbytes)
    orc.wide4_check(wn, nodes, 4)
    rng = np.random.default_rng(1)
    n = 20000
    rays = np.zeros((n, 8), dtype=np.float32)
    rays[:, 0] = rng.uniform(-0.5, 5.5, n); rays[:, 1] = rng.uniform(-0.5, 1.5, n); rays[:, 2] = 20.0
    rays[:, 4:7] = (0, 0, -1)
    rays[: n // 2, 4:7] += rng.normal(0, 0.05, (n // 2, 3)).astype(np.float32)
    rays[:, 7] = 100.0
    ref = orc.compact_trace(nodes, woop, idx, rays, True)
    got = orc.wide4_trace(wn, woop, idx, rays, True)
    assert np.array_equal(ref[:, :2], got[:, :2])
