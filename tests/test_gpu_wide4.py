"""GPU parity of the Wide4 and multi-ray-lane kernels (csrc/nt_wide.cu), the resident-BVH generation rule and the
traversal-stack guard.  Bars: Wide4 kernels are bit-identical to the oracle's emulation of the product's Wide4 traversal
(ids, t bits, any-hit included), which tests/test_wide4.py ties to the reference's Compact tracer; the "mr" kernels over the
binary nodes are bit-identical to the one-ray binary kernel (same per-ray visiting order)."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame(gpu_host, orc):
    verts, tris = scenes.room(30_000, seed=13, wall_frac=0.3)
    cam = camera.named_camera("conference")
    scene = gpu_host.Scene(verts, tris)
    cpu = orc.CpuBVH(verts, tris, orc.BUILDER_SPLIT, 1, 1)
    nodes, woop, idx = cpu.compact()
    w, h = 256, 192
    prim = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(gpu_host.CudaBVH(nodes, woop, idx))
    tracer.traceBatch(prim)
    batches = [("primary", prim, True)]
    rg = gpu_host.RayGen(1 << 18)
    for name, dist, closest in (("AO", 5.0, False), ("diffuse", cam.far, True)):
        rb = gpu_host.RayBuffer()
        ok, _ = rg.ao(rb, prim, scene, 16, dist, True, gpu_host.FIXED_AO_SEED)
        assert ok
        rb.setNeedClosestHit(closest)
        batches.append((name, rb, closest))
    return dict(verts=verts, tris=tris, scene=scene, nodes=nodes, woop=woop, idx=idx, batches=batches)


def _trace(gpu_host, frame, kernel, bvh):
    tracer = gpu_host.CudaBVHTracer()
    tracer.setKernel(kernel)
    tracer.setBVH(bvh)
    out = {}
    for name, rb, closest in frame["batches"]:
        rb.setNeedClosestHit(closest)
        assert tracer.traceBatch(rb) > 0.0
        out[name] = rb.results_host().copy()
    tracer.setKernel("b200_persistent_speculative_while_while")
    return out


@pytest.mark.parametrize("kernel", ["b200_wide4", "b200_wide4_mr", "b200_wide4_sw", "b200_wide4_fastmath"])
def test_wide4_kernels_are_bit_identical_to_the_emulation(gpu_host, orc, frame, kernel):
    bvh = gpu_host.CudaBVH(frame["nodes"], frame["woop"], frame["idx"])
    got = _trace(gpu_host, frame, kernel, bvh)
    wn, _ = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, frame["nodes"], frame["woop"].nbytes)
    for name, rb, closest in frame["batches"]:
        rays = rb.rays_host()
        ref = orc.wide4_trace(wn, frame["woop"], frame["idx"], rays, closest)
        if "fastmath" in kernel:
            # the -use_fast_math triangle arithmetic of the reference's GPU kernels: same hits up to rounding
            assert (got[name][:, 0] == ref[:, 0]).mean() >= 0.995 if not closest else (got[name][:, 0] == ref[:, 0]).mean() >= 0.9995
            continue
        assert np.array_equal(got[name][:, 0], ref[:, 0]), name
        assert np.array_equal(got[name][:, 1], ref[:, 1]), name
        # and against the reference's binary Compact tracer: north_star tolerances
        flat = orc.compact_trace(frame["nodes"], frame["woop"], frame["idx"], rays, closest)
        assert ((got[name][:, 0] >= 0) == (flat[:, 0] >= 0)).mean() >= 0.9999
        if closest:
            same = got[name][:, 0] == flat[:, 0]
            assert same.mean() >= 0.9999
            hit = same & (flat[:, 0] >= 0)
            assert np.array_equal(got[name][hit, 1], flat[hit, 1])


@pytest.mark.parametrize("kernel", ["b200_mr", "b200_mr_fastmath", "b200_sw", "b200_sw_fastmath"])
def test_multi_ray_lane_kernel_equals_the_one_ray_kernel(gpu_host, frame, kernel):
    bvh = gpu_host.CudaBVH(frame["nodes"], frame["woop"], frame["idx"])
    base = "b200_persistent_speculative_while_while" + ("_fastmath" if "fastmath" in kernel else "")
    a = _trace(gpu_host, frame, base, bvh)
    b = _trace(gpu_host, frame, kernel, bvh)
    for name in a:
        assert np.array_equal(a[name], b[name]), name          # id, t, u, v: every bit, any-hit included


def _canonical_wide(wn):
    """numbering-independent form of a Wide4 tree: node payloads in depth-first slot order, inner links replaced by a marker"""
    wn = wn.reshape(-1, 16)
    out, stack = [], [0]
    while stack:
        w = wn[stack.pop()]
        links = w[12:16].view(np.int32)
        out.append(w[:12].tobytes() + np.where(links < 0, links, 0).astype(np.int32).tobytes())
        stack.extend(int(l) for l in links[::-1] if l >= 0)
    return out


@pytest.mark.parametrize("layout", [capi.LAYOUT_COMPACT, capi.LAYOUT_COMPACT2])
def test_device_conversion_equals_the_host_statement_of_the_format(gpu_host, frame, layout):
    # what the library traces (wide4_convert_kernel, derived on the device) against nt_bvh_wide4_convert_host on the same tree:
    # same nodes, plane bytes, child slots and leaves; only the numbering may differ
    capi.bvh_set_build_layout(layout)
    try:
        bvh = gpu_host.HLBVHBuilder(frame["scene"], gpu_host.HLBVHParams(True, 4, 8, 0.001))
        nodes, woop, idx, lay = capi.bvh_download()
        assert lay == layout
        dev, dev_depth = capi.bvh_wide4_download()
        ref, ref_depth = capi.bvh_wide4_convert_host(layout, nodes, woop.nbytes)
    finally:
        capi.bvh_set_build_layout(capi.LAYOUT_COMPACT)
    assert dev.size == ref.size and dev_depth == ref_depth
    assert _canonical_wide(dev) == _canonical_wide(ref)
    del bvh


def test_wide4_on_a_gpu_built_compact2_tree(gpu_host, orc, frame):
    capi.bvh_set_build_layout(capi.LAYOUT_COMPACT2)
    try:
        bvh = gpu_host.HLBVHBuilder(frame["scene"], gpu_host.HLBVHParams(True, 4, 8, 0.001))
        assert bvh.getLayout() == capi.LAYOUT_COMPACT2           # the handle reports what the library built
        got = _trace(gpu_host, frame, "b200_wide4_compact2", bvh)
        ref = _trace(gpu_host, frame, "b200_persistent_speculative_while_while_compact2", bvh)
    finally:
        capi.bvh_set_build_layout(capi.LAYOUT_COMPACT)
    for name, rb, closest in frame["batches"]:
        assert ((got[name][:, 0] >= 0) == (ref[name][:, 0] >= 0)).mean() >= 0.9999
        if closest:
            same = got[name][:, 0] == ref[name][:, 0]
            assert same.mean() >= 0.9999
            assert np.array_equal(got[name][same, 1], ref[name][same, 1])


def test_resident_handles_follow_the_generation(gpu_host, frame):
    scene = frame["scene"]
    a = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(False, 10, 8, 0.001))
    gen_a = capi.bvh_generation()
    assert a.generation == gen_a and gen_a > 0
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(a)
    name, rb, closest = frame["batches"][0]
    rb.setNeedClosestHit(True)
    tracer.traceBatch(rb)
    res_a = rb.results_host().copy()
    host_copy = (a.getNodeBuffer().copy(), a.getTriWoopBuffer().copy(), a.getTriIndexBuffer().copy())     # materialises the handle
    b = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(False, 10, 1, 0.001))                              # replaces the resident BVH
    assert capi.bvh_generation() != gen_a
    # the stale handle has a host copy: tracing through it uploads it again instead of tracing b's tree
    tracer.traceBatch(rb)
    assert np.array_equal(rb.results_host(), res_a)
    assert a.generation == capi.bvh_generation()
    # a device-only handle that was replaced is refused, not silently swapped
    c = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(False, 10, 4, 0.001))
    assert b.nodes is None                                        # b has no host copy and is no longer resident
    with pytest.raises(capi.NtError, match="replaced"):
        tracer.setBVH(b)
    tracer.setBVH(c)
    del host_copy


def test_traversal_stack_overflow_is_reported_not_silent(gpu_host):
    depth = 200                                                   # a chain deeper than the 96-entry traversal stack
    nodes = np.zeros((depth, 16), dtype=np.int32)
    f = nodes.view(np.float32)
    f[:, 0:12:2] = 0.0
    f[:, 1:12:2] = 1.0
    for i in range(depth):
        nodes[i, 12] = (i + 1) * 64 if i + 1 < depth else ~0     # child 0: the next node (entered first), child 1: an empty leaf (pushed)
        nodes[i, 13] = ~0
    woop = np.full(4, np.int32(-2**31), dtype=np.int32)
    idx = np.zeros(1, dtype=np.int32)
    capi.set_kernel("b200_persistent_speculative_while_while")
    capi.bvh_upload(capi.LAYOUT_COMPACT, nodes.reshape(-1), woop, idx)
    rays = np.zeros((64, 8), dtype=np.float32)
    rays[:, 0:3] = (0.5, 0.5, -1.0)
    rays[:, 4:7] = (0.0, 0.0, 1.0)
    rays[:, 7] = 10.0
    res = np.zeros((64, 4), dtype=np.int32)
    with pytest.raises(capi.NtError, match="too deep"):
        capi.trace_batch(rays, res, 64, True)
    # the flag is consumed: a shallow tree traces normally afterwards
    capi.bvh_upload(capi.LAYOUT_COMPACT, nodes[-1:].reshape(-1).copy() * 0 + nodes[-1].reshape(-1), woop, idx)
    capi.trace_batch(rays, res, 64, True)
    assert (res[:, 0] == -1).all()
    # the Wide4 conversion refuses such a tree before tracing
    capi.bvh_upload(capi.LAYOUT_COMPACT, nodes.reshape(-1), woop, idx)
    capi.set_kernel("b200_wide4")
    with pytest.raises(capi.NtError, match="too deep"):
        capi.trace_batch(rays, res, 64, True)
    capi.set_kernel("b200_persistent_speculative_while_while")


@pytest.mark.timeout(120)
def test_device_conversion_refuses_malformed_trees(gpu_host):
    # the same malformed inputs tests/test_wide4.py gives the host routine, through the conversion the library itself runs (on the device):
    # a cycle must end in an error, not in a kernel that waits for ever
    nodes = np.zeros((4, 16), dtype=np.int32)
    f = nodes.view(np.float32)
    f[:, 0:12:2] = 0.0
    f[:, 1:12:2] = 1.0
    for i in range(4):
        nodes[i, 12] = (i + 1) * 64 if i + 1 < 4 else ~0
        nodes[i, 13] = ~0
    woop = np.full(4, np.int32(-2**31), dtype=np.int32)
    idx = np.zeros(1, dtype=np.int32)
    rays = np.zeros((64, 8), dtype=np.float32)
    rays[:, 0:3] = (0.5, 0.5, -1.0)
    rays[:, 4:7] = (0.0, 0.0, 1.0)
    rays[:, 7] = 10.0
    res = np.zeros((64, 4), dtype=np.int32)
    try:
        for word, value, pattern in ((3, 64, "cycle|outside"), (2, 4096, "outside"), (3, ~50, "outside")):
            bad = nodes.copy()
            bad[word, 12] = value
            capi.set_kernel("b200_persistent_speculative_while_while")
            capi.bvh_upload(capi.LAYOUT_COMPACT, bad.reshape(-1), woop, idx)
            capi.set_kernel("b200_wide4")
            with pytest.raises(capi.NtError, match=pattern):
                capi.trace_batch(rays, res, 64, True)
        # and the well-formed chain converts and traces (every ray misses: the leaves are empty)
        capi.set_kernel("b200_persistent_speculative_while_while")
        capi.bvh_upload(capi.LAYOUT_COMPACT, nodes.reshape(-1), woop, idx)
        capi.set_kernel("b200_wide4")
        capi.trace_batch(rays, res, 64, True)
        assert (res[:, 0] == -1).all()
        wn, depth = capi.bvh_wide4_download()
        ref, ref_depth = capi.bvh_wide4_convert_host(capi.LAYOUT_COMPACT, nodes.reshape(-1), woop.nbytes)
        assert depth == ref_depth and _canonical_wide(wn) == _canonical_wide(ref)
    finally:
        capi.set_kernel("b200_persistent_speculative_while_while")
