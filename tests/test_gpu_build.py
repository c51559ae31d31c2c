"""GPU parity of the LBVH builder against the CPU restatement of the reference HLBVHBuilder:
Morton codes and sort order bit-exact, the (start, split, end) range tree identical (compared in the
numbering-independent canonical form of SURVEY.md App. B-5), child boxes and Woop rows equal, SAH equal."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu


def _gpu_build(gpu_host, verts, tris, lo, hi, leaf=8, eps=0.001):
    sec = capi.bvh_build(capi.BUILDER_LBVH, np.ascontiguousarray(verts, np.float32), np.ascontiguousarray(tris, np.int32), lo, hi, 10, leaf, eps)
    nodes, woop, idx, layout = capi.bvh_download()
    assert layout == capi.LAYOUT_COMPACT
    keys, order = capi.bvh_build_debug(len(tris))
    return nodes, woop, idx, keys, order, sec


def _assert_same_tree(orc, gpu, ref):
    gn, gw, gi = gpu
    cg = orc.canonical(gn, gw, gi)
    cr = orc.canonical(ref.nodes, ref.woop, ref.tri_index)
    assert np.array_equal(cg.inner, cr.inner), "range tree differs"
    assert np.array_equal(cg.leaf_sizes, cr.leaf_sizes)
    assert np.array_equal(cg.tris, cr.tris)
    assert np.array_equal(cg.boxes, cr.boxes), "child boxes differ"          # value equality (-0 == +0)
    # bit-exact, except that NaN payloads (degenerate zero-area triangles: 1/0 * 0) differ between x86 and the GPU
    wg, wr = cg.woop.view(np.uint32), cr.woop.view(np.uint32)
    assert ((wg == wr) | (np.isnan(cg.woop) & np.isnan(cr.woop))).all(), "Woop rows differ"
    assert len(gn) == len(ref.nodes) and len(gw) == len(ref.woop) and len(gi) == len(ref.tri_index)


def _cells_scene(codes_xyz, reps):
    """Tiny triangles centred in grid cells of a [0,1024]^3 box -> exact, chosen Morton codes."""
    verts, tris = [], []
    for (x, y, z), r in zip(codes_xyz, reps):
        for k in range(r):
            c = np.array([x + 0.5, y + 0.5, z + 0.5], np.float32)
            d = np.float32(0.01 + 0.0001 * (k % 7))
            base = len(verts)
            verts += [c + (-d, -d, 0), c + (d, -d, 0), c + (0, d, 0)]
            tris.append((base, base + 1, base + 2))
    return np.array(verts, np.float32), np.array(tris, np.int32), np.zeros(3, np.float32), np.full(3, 1024.0, np.float32)


def _demorton(code):
    def compact(v):
        r = 0
        for i in range(10):
            r |= ((v >> (3 * i)) & 1) << i
        return r
    return compact(code), compact(code >> 1), compact(code >> 2)


@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_lbvh_matches_reference_restatement(gpu_host, orc, leaf):
    verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
    lo, hi = scenes.bbox(verts)
    nodes, woop, idx, keys, order, sec = _gpu_build(gpu_host, verts, tris, lo, hi, leaf)
    ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf, epsilon=0.001)
    assert np.array_equal(np.sort(orc.morton(verts, tris, lo, hi)), keys)
    assert np.array_equal(keys, ref.sorted_keys), "Morton codes / sort differ"
    assert np.array_equal(order, ref.sorted_idx), "sort is not stable"
    _assert_same_tree(orc, (nodes, woop, idx), ref)
    sg, sr = orc.compact_sah(nodes, woop), orc.compact_sah(ref.nodes, ref.woop)
    assert abs(sg["sah"] - sr["sah"]) <= 0.005 * sr["sah"]
    assert sg["num_tris"] == len(tris)


def test_lbvh_clustered_soup_with_duplicates(gpu_host, orc):
    verts, tris = scenes.soup_uniform(60_000, seed=9, clustered=True)
    lo, hi = scenes.bbox(verts)
    # shrink precision so that many triangles share a cell: quantise against a much larger box
    lo2, hi2 = lo - 40.0, hi + 40.0
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, verts, tris, lo2, hi2, 4)
    ref = orc.lbvh_build(verts, tris, lo2, hi2, hlbvh=False, leaf_size=4)
    assert (np.diff(keys.astype(np.int64)) == 0).sum() > 1000, "test needs duplicate codes"
    assert np.array_equal(keys, ref.sorted_keys) and np.array_equal(order, ref.sorted_idx)
    _assert_same_tree(orc, (nodes, woop, idx), ref)


@pytest.mark.parametrize("n,leaf", [(1, 8), (2, 8), (5, 8), (9, 8), (3, 1), (64, 1)])
def test_lbvh_tiny_inputs(gpu_host, orc, n, leaf):
    verts, tris = scenes.soup_uniform(n, seed=3)
    lo, hi = scenes.bbox(verts)
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, verts, tris, lo, hi, leaf)
    ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=leaf)
    _assert_same_tree(orc, (nodes, woop, idx), ref)


@pytest.mark.parametrize("n", [2047, 2048, 2049, 4096, 6145])
def test_lbvh_sizes_around_the_sort_tile(gpu_host, orc, n):
    # the radix scatter stages tiles of 2048 keys in shared memory: a partly filled last tile, exactly full tiles, one key over
    verts, tris = scenes.soup_uniform(n, seed=11, clustered=(n % 2 == 0))
    lo, hi = scenes.bbox(verts)
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, verts, tris, lo, hi, 4)
    ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=4)
    assert np.array_equal(keys, ref.sorted_keys) and np.array_equal(order, ref.sorted_idx)
    _assert_same_tree(orc, (nodes, woop, idx), ref)


def test_lbvh_full_tiles_of_one_digit(gpu_host, orc):
    # 5000 triangles in one Morton cell: every key of two full tiles lands in the same digit bucket in all four passes
    v, t, lo, hi = _cells_scene([(5, 6, 7), (900, 3, 512)], [5000, 300])
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, v, t, lo, hi, 8)
    ref = orc.lbvh_build(v, t, lo, hi, hlbvh=False, leaf_size=8)
    assert np.array_equal(keys, ref.sorted_keys) and np.array_equal(order, ref.sorted_idx), "sort is not stable"
    _assert_same_tree(orc, (nodes, woop, idx), ref)


def test_lbvh_all_identical_keys_median_rule(gpu_host, orc):
    v, t, lo, hi = _cells_scene([(5, 6, 7)], [1000])
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, v, t, lo, hi, 8)
    assert len(np.unique(keys)) == 1
    ref = orc.lbvh_build(v, t, lo, hi, hlbvh=False, leaf_size=8)
    _assert_same_tree(orc, (nodes, woop, idx), ref)
    c = orc.canonical(nodes, woop, idx)
    assert (c.inner[:, 2] == -1).all()              # `level % 3` with level == -1


def test_lbvh_forced_leaves_29_levels_down(gpu_host, orc):
    # keys 0 (x M), 1, 2, 4, ..., 2^29: a radix chain 29 levels deep ending in a long duplicate run
    cells = [_demorton(0)] + [_demorton(1 << j) for j in range(30)]
    for m, leaf in [(100, 8), (37, 1)]:
        v, t, lo, hi = _cells_scene(cells, [m] + [1] * 30)
        nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, v, t, lo, hi, leaf)
        assert keys[m] == 1 and keys[-1] == (1 << 29)
        ref = orc.lbvh_build(v, t, lo, hi, hlbvh=False, leaf_size=leaf)
        _assert_same_tree(orc, (nodes, woop, idx), ref)
        assert orc.canonical(nodes, woop, idx).leaf_sizes.max() == m     # the forced leaf holds the whole run
    # run entered 20 levels down, bisected for 9 more levels, then forced
    cells = [_demorton(0)] + [_demorton(1 << j) for j in range(10, 30)]
    v, t, lo, hi = _cells_scene(cells, [10_000] + [1] * 20)
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, v, t, lo, hi, 4)
    ref = orc.lbvh_build(v, t, lo, hi, hlbvh=False, leaf_size=4)
    _assert_same_tree(orc, (nodes, woop, idx), ref)
    assert orc.canonical(nodes, woop, idx).leaf_sizes.max() > 4


def test_lbvh_planar_scene_zero_extent_axis(gpu_host, orc):
    # all z equal: step.z == 0, the quantiser sees 0/0 (reference: calcMorton has no guard)
    rng = np.random.default_rng(5)
    n = 3000
    c = rng.uniform(0, 10, size=(n, 2)).astype(np.float32)
    verts = np.zeros((n * 3, 3), np.float32)
    verts[0::3, :2] = c; verts[1::3, :2] = c + (0.1, 0); verts[2::3, :2] = c + (0, 0.1)
    verts[:, 2] = 2.5
    tris = np.arange(n * 3, dtype=np.int32).reshape(n, 3)
    lo, hi = scenes.bbox(verts)
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, verts, tris, lo, hi, 8)
    ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=8)
    assert np.array_equal(keys, ref.sorted_keys)
    _assert_same_tree(orc, (nodes, woop, idx), ref)


def test_gpu_built_bvh_traces_like_cpu(gpu_host, orc):
    verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
    scene = gpu_host.Scene(verts, tris)
    bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(False, 10, 8, 0.001))
    assert bvh.getGPUTime() > 0
    cam = camera.named_camera("conference")
    w, h = 256, 192
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    rays = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    tracer.traceBatch(rays)
    got = rays.results_host()
    ref = orc.compact_trace(bvh.getNodeBuffer(), bvh.getTriWoopBuffer(), bvh.getTriIndexBuffer(), rays.rays_host(), True)
    assert (got[:, 0] == ref[:, 0]).mean() >= 0.9999
    brute = orc.brute_trace(verts, tris, rays.rays_host()[:4096], True)
    assert (got[:4096, 0] == brute[:, 0]).mean() >= 0.999


@pytest.mark.parametrize("bits,leaf,ntris", [(4, 8, 20_000), (2, 8, 20_000), (4, 1, 6_000), (6, 4, 60_000), (3, 8, 2_000)])
def test_hlbvh_matches_reference_restatement(gpu_host, orc, bits, leaf, ntris):
    """HLBVH: clusters + top-level binned SAH + per-cluster LBVH; the restatement runs the reference kernels with a
    serial schedule and the GPU stage is deterministic with the same tie rules, so the trees must be identical."""
    verts, tris = scenes.room(ntris, seed=7, wall_frac=0.3)
    lo, hi = scenes.bbox(verts)
    capi.bvh_build(capi.BUILDER_HLBVH, verts, tris, lo, hi, bits, leaf, 0.001)
    nodes, woop, idx, layout = capi.bvh_download()
    ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=True, hlbvh_bits=bits, leaf_size=leaf, epsilon=0.001)
    assert ref.num_clusters > 1
    sg, sr = orc.compact_sah(nodes, woop), orc.compact_sah(ref.nodes, ref.woop)
    assert abs(sg["sah"] - sr["sah"]) <= 0.005 * sr["sah"], (sg, sr)          # north-star bar
    assert sg["num_tris"] == len(tris)
    _assert_same_tree(orc, (nodes, woop, idx), ref)


def test_hlbvh_soup_and_fairy_scene(gpu_host, orc):
    for verts, tris in (scenes.soup_uniform(50_000, seed=4), scenes.teapot_in_stadium(40_000, seed=3)):
        lo, hi = scenes.bbox(verts)
        capi.bvh_build(capi.BUILDER_HLBVH, verts, tris, lo, hi, 4, 8, 0.001)
        nodes, woop, idx, _ = capi.bvh_download()
        ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=True, hlbvh_bits=4, leaf_size=8)
        sg, sr = orc.compact_sah(nodes, woop), orc.compact_sah(ref.nodes, ref.woop)
        assert abs(sg["sah"] - sr["sah"]) <= 0.005 * sr["sah"]
        c = orc.canonical(nodes, woop, idx)
        assert sorted(c.tris.tolist()) == list(range(len(tris)))
        _assert_same_tree(orc, (nodes, woop, idx), ref)


def test_hlbvh_traces_like_cpu_and_bits10_is_lbvh(gpu_host, orc):
    verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
    scene = gpu_host.Scene(verts, tris)
    bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(True, 4, 8, 0.001))        # Renderer.cpp:201-209 defaults
    cam = camera.named_camera("conference")
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    rays = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, 256, 192), 256, 192, cam.far)
    tracer.traceBatch(rays)
    got = rays.results_host()
    ref = orc.compact_trace(bvh.getNodeBuffer(), bvh.getTriWoopBuffer(), bvh.getTriIndexBuffer(), rays.rays_host(), True)
    assert (got[:, 0] == ref[:, 0]).mean() >= 0.9999
    brute = orc.brute_trace(verts, tris, rays.rays_host()[:4096], True)
    assert (got[:4096, 0] == brute[:, 0]).mean() >= 0.999
    # hlbvhBits == 10 selects the plain LBVH path (HLBVHBuilder.cpp:44-47)
    lo, hi = scenes.bbox(verts)
    capi.bvh_build(capi.BUILDER_HLBVH, verts, tris, lo, hi, 10, 8, 0.001)
    a = capi.bvh_download()
    capi.bvh_build(capi.BUILDER_LBVH, verts, tris, lo, hi, 10, 8, 0.001)
    b = capi.bvh_download()
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
    from ntrace_b200 import NtError
    with pytest.raises(NtError, match="hlbvhBits"):
        capi.bvh_build(capi.BUILDER_HLBVH, verts, tris, lo, hi, 0, 8, 0.001)


@pytest.mark.parametrize("builder,bits", [(capi.BUILDER_LBVH, 10), (capi.BUILDER_HLBVH, 4)])
def test_sah_collapse_mode_is_valid_and_not_worse(gpu_host, orc, builder, bits):
    """collapse=sah (not part of the reference): every triangle exactly once, leaves <= maxLeaf, SAH not above the
    count-rule tree, traversal results identical to the count-rule tree (closest hit is tree-independent up to ties)."""
    verts, tris = scenes.room(30_000, seed=7, wall_frac=0.3)
    lo, hi = scenes.bbox(verts)
    capi.bvh_set_collapse(0)
    capi.bvh_build(builder, verts, tris, lo, hi, bits, 8, 0.001)
    base = capi.bvh_download()
    try:
        capi.bvh_set_collapse(1, 8)
        capi.bvh_build(builder, verts, tris, lo, hi, bits, 8, 0.001)
        coll = capi.bvh_download()
    finally:
        capi.bvh_set_collapse(0)
    cb, cc = orc.canonical(*base[:3]), orc.canonical(*coll[:3])
    assert sorted(cc.tris.tolist()) == list(range(len(tris)))
    assert cc.leaf_sizes.max() <= 8 and cc.leaf_sizes.min() >= 1
    sb, sc = orc.compact_sah(base[0], base[1])["sah"], orc.compact_sah(coll[0], coll[1])["sah"]
    assert sc <= sb * 1.0001, (sc, sb)
    assert len(coll[0]) != len(base[0])                                   # the tree really changed
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    a = orc.compact_trace(*base[:3], rays, True)
    b = orc.compact_trace(*coll[:3], rays, True)
    same = a[:, 0] == b[:, 0]
    rel = np.abs(a[:, 1].view(np.float32) - b[:, 1].view(np.float32)) / np.maximum(np.abs(a[:, 1].view(np.float32)), 1e-30)
    assert same.mean() >= 0.999 and ((~same) & (rel > 1e-4)).sum() == 0
    # and the CUDA kernel traverses it like the CPU does
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(gpu_host.CudaBVH(*coll[:3]))
    rb = gpu_host.RayBuffer(); rb.setRays(rays)
    tracer.traceBatch(rb)
    assert np.array_equal(rb.results_host()[:, 0], b[:, 0])


def test_randomised_small_inputs_match_restatement(gpu_host, orc):
    """Sweep of small random scenes (sizes, leaf sizes, duplicate density, both builders): GPU tree == restated tree."""
    rng = np.random.default_rng(1234)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        leaf = int(rng.choice([1, 2, 3, 8]))
        verts, tris = scenes.soup_uniform(n, seed=100 + trial, clustered=bool(trial % 3 == 0))
        lo, hi = scenes.bbox(verts)
        pad = np.float32(rng.choice([0.0, 4.0, 60.0]))              # coarser grid -> more duplicate codes
        lo, hi = lo - pad, hi + pad
        hl = bool(trial % 2) and n > 4
        bits = int(rng.choice([2, 4, 7]))
        capi.bvh_build(capi.BUILDER_HLBVH if hl else capi.BUILDER_LBVH, verts, tris, lo, hi, bits if hl else 10, leaf, 0.001)
        nodes, woop, idx, _ = capi.bvh_download()
        ref = orc.lbvh_build(verts, tris, lo, hi, hlbvh=hl, hlbvh_bits=bits, leaf_size=leaf, epsilon=0.001)
        if hl and ref.num_clusters < 2:
            continue                                                 # single-cell scenes: the reference is undefined there
        try:
            _assert_same_tree(orc, (nodes, woop, idx), ref)
        except AssertionError as e:
            raise AssertionError(f"trial {trial}: n={n} leaf={leaf} hlbvh={hl} bits={bits} pad={pad}: {e}") from None


def test_trace_through_a_30_level_chain(gpu_host, orc):
    """Deepest LBVH shape (one split per Morton bit, then a forced leaf): the traversal stack spills past its 8
    shared-memory entries into the local array; results must still match the CPU trace of the same buffers."""
    cells = [_demorton(0)] + [_demorton(1 << j) for j in range(30)]
    v, t, lo, hi = _cells_scene(cells, [60] + [3] * 30)
    nodes, woop, idx, keys, order, _ = _gpu_build(gpu_host, v, t, lo, hi, 1)
    assert orc.compact_sah(nodes, woop)["max_depth"] >= 30
    rng = np.random.default_rng(3)
    n = 4096
    rays = np.zeros((n, 8), np.float32)
    pick = t[rng.integers(0, len(t), n)]
    targets = (v[pick[:, 0]] + v[pick[:, 1]] + v[pick[:, 2]]) / np.float32(3.0) + rng.normal(0, 0.002, (n, 3)).astype(np.float32)
    origins = np.array([1500.0, 1400.0, 1300.0], np.float32) + rng.normal(0, 50, (n, 3)).astype(np.float32)
    d = targets - origins
    rays[:, 0:3] = origins; rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True); rays[:, 7] = 1e5
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(gpu_host.CudaBVH(nodes, woop, idx))
    for kernel in ("b200_persistent_speculative_while_while", "b200_speculative_while_while"):
        tracer.setKernel(kernel)
        rb = gpu_host.RayBuffer(); rb.setRays(rays)
        tracer.traceBatch(rb)
        got = rb.results_host()
        ref = orc.compact_trace(nodes, woop, idx, rays, True)
        assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, 1], ref[:, 1])
        assert (ref[:, 0] >= 0).mean() > 0.3
    tracer.setKernel("b200_persistent_speculative_while_while")


def test_builder_emits_compact2_on_request(gpu_host, orc):
    """nt_bvh_set_build_layout(Compact2): same tree, inner links are offsets / 16 (what lifts the 1.98 GB node limit of Compact)."""
    from ntrace_b200 import capi
    verts, tris = scenes.room(40_000, seed=17, wall_frac=0.3)
    lo, hi = scenes.bbox(verts)
    scene = gpu_host.Scene(verts, tris)
    cam = camera.named_camera("conference")
    rays, _, _ = orc.raygen_primary(cam.position, camera.nscreen_to_world(cam, 256, 192), 256, 192, cam.far)
    out = {}
    try:
        for layout, kernel in ((4, "b200_persistent_speculative_while_while"), (5, "b200_persistent_speculative_while_while_compact2")):
            for builder, bits in ((capi.BUILDER_LBVH, 10), (capi.BUILDER_HLBVH, 4)):
                capi.bvh_set_build_layout(layout)
                capi.bvh_build(builder, scene.vtxPos, scene.triVtxIndex, lo, hi, bits, 8, 0.001)
                nodes, woop, idx, got_layout = capi.bvh_download()
                assert got_layout == layout
                tracer = gpu_host.CudaBVHTracer(); tracer.setKernel(kernel)
                bvh = gpu_host.CudaBVH(layout=layout); bvh.resident = True
                tracer.setBVH(bvh)
                rb = gpu_host.RayBuffer(); rb.setRays(rays)
                tracer.traceBatch(rb)
                out[(layout, builder)] = (nodes.reshape(-1, 16), woop, idx, rb.results_host()[:, :2].copy())
    finally:
        capi.bvh_set_build_layout(4)
    for builder in (capi.BUILDER_LBVH, capi.BUILDER_HLBVH):
        n4, w4, i4, r4 = out[(4, builder)]
        n5, w5, i5, r5 = out[(5, builder)]
        assert np.array_equal(w4, w5) and np.array_equal(i4, i5) and np.array_equal(r4, r5)
        assert np.array_equal(n4[:, :12], n5[:, :12]) and np.array_equal(n4[:, 14:], n5[:, 14:])
        assert np.array_equal(np.where(n4[:, 12:14] >= 0, n4[:, 12:14] // 16, n4[:, 12:14]), n5[:, 12:14])
    with pytest.raises(capi.NtError, match="Compact"):
        capi.bvh_set_build_layout(0)


@pytest.mark.parametrize("bits,collapse", [(10, 0), (4, 0), (2, 1)])
def test_device_sah_metric_matches_the_reference_formula(gpu_host, orc, bits, collapse):
    """nt_bvh_sah (one parallel pass on the device) against the restated BVHNode::computeSubtreeProbabilities walk of the same tree."""
    verts, tris = scenes.room(40_000, seed=17, wall_frac=0.3)
    lo, hi = scenes.bbox(verts)
    capi.bvh_set_collapse(collapse, 8)
    try:
        capi.bvh_build(capi.BUILDER_HLBVH, np.ascontiguousarray(verts, np.float32), np.ascontiguousarray(tris, np.int32), lo, hi, bits, 8, 0.001)
    finally:
        capi.bvh_set_collapse(0, 0)
    got = capi.bvh_sah()
    nodes, woop, idx, _ = capi.bvh_download()
    ref = orc.compact_sah(nodes, woop)
    assert got["num_inner"] == ref["num_inner"] and got["num_leaf"] == ref["num_leaf"] and got["num_tris"] == ref["num_tris"] == len(tris)
    assert abs(got["sah"] - ref["sah"]) <= 2e-4 * ref["sah"], (got["sah"], ref["sah"])     # fp32 probability products there, fp64 areas here
    # an uploaded CPU-built SplitBVH too (different leaf shape: one triangle per leaf, duplicated references)
    v2, t2 = scenes.room(9_000, seed=3, wall_frac=0.3)
    cpu = orc.CpuBVH(v2, t2, orc.BUILDER_SPLIT, 1, 1)
    n2, w2, i2 = cpu.compact()
    capi.bvh_upload(capi.LAYOUT_COMPACT, n2, w2, i2)
    got2, ref2 = capi.bvh_sah(), orc.compact_sah(n2, w2)
    assert got2["num_leaf"] == ref2["num_leaf"] and abs(got2["sah"] - ref2["sah"]) <= 2e-4 * ref2["sah"]
