"""The HEADLINE configuration under test, not a scaled-down stand-in: BASELINE.json configs[1] exactly as bench.py times it
(Conference stand-in room(283000, seed=2), 1024x768, 32 spp).  Two trees: the reference Renderer's (HLBVH, hlbvhBits 4, leaf 8,
count rule — Renderer.cpp:201-209), whose whole structure is compared with the restated reference builder, and the bench's
(hlbvhBits 2 + SAH-guided collapse), which has no reference counterpart and is therefore checked through the traced results
only.  Rays: the full primary batch and one full 1 Mi-ray batch of each secondary type from the middle of the frame, every ray
compared with the oracle's flat tracer on the BVH downloaded from the device (north_star tolerances)."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu

W, H, SPP = 1024, 768, 32


@pytest.fixture(scope="module")
def conference(gpu_host):
    verts, tris, cam_name = scenes.config_scene("conference")
    return verts, tris, camera.named_camera(cam_name), gpu_host.Scene(verts, tris)


def _frame_batches(gpu_host, tracer, scene, cam):
    prim = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far)
    tracer.traceBatch(prim)
    out = [("primary", prim, True)]
    for name, dist, closest in (("AO", 5.0, False), ("diffuse", cam.far, True)):
        rg = gpu_host.RayGen(1 << 20)
        new, rb = True, None
        for _ in range(12):                                       # the 12th batch: the middle of the image
            rb = gpu_host.RayBuffer()
            ok, new = rg.ao(rb, prim, scene, SPP, dist, new, gpu_host.FIXED_AO_SEED)
            assert ok
        rb.setNeedClosestHit(closest)
        out.append((name, rb, closest))
    return out


def _check_against_oracle(orc, tracer, batches, nodes, woop, idx):
    for name, rb, closest in batches:
        assert tracer.traceBatch(rb) > 0.0
        got, rays = rb.results_host(), rb.rays_host()
        ref = orc.compact_trace(nodes, woop, idx, rays, closest)
        hit_g, hit_r = got[:, 0] >= 0, ref[:, 0] >= 0
        assert (hit_g == hit_r).mean() >= 0.9999, name
        if closest:
            same = got[:, 0] == ref[:, 0]
            assert same.mean() >= 0.9999, (name, same.mean())
            tg, tr = got[:, 1].view(np.float32), ref[:, 1].view(np.float32)
            both = hit_g & hit_r
            rel = np.abs(tg - tr)[both] / np.maximum(np.abs(tr[both]), 1e-30)
            assert rel[same[both]].max() <= 1e-5, name
            assert rel.max() <= 1e-4, name                      # id mismatches only where t agrees


def test_reference_renderer_tree_at_full_size(gpu_host, orc, conference):
    verts, tris, cam, scene = conference
    capi.bvh_set_collapse(0, 0)
    bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(True, 4, 8, 0.001))          # Renderer.cpp:201-209
    nodes, woop, idx = bvh.getNodeBuffer(), bvh.getTriWoopBuffer(), bvh.getTriIndexBuffer()
    ref = orc.lbvh_build(verts, tris, scene.bboxMin, scene.bboxMax, hlbvh=True, hlbvh_bits=4, leaf_size=8, epsilon=0.001)
    cg, cr = orc.canonical(nodes, woop, idx), orc.canonical(ref.nodes, ref.woop, ref.tri_index)
    assert np.array_equal(cg.inner, cr.inner) and np.array_equal(cg.leaf_sizes, cr.leaf_sizes) and np.array_equal(cg.tris, cr.tris)
    assert np.array_equal(cg.boxes, cr.boxes)
    wg, wr = cg.woop.view(np.uint32), cr.woop.view(np.uint32)                 # bit-exact (NaN payloads of degenerate triangles aside)
    assert ((wg == wr) | (np.isnan(cg.woop) & np.isnan(cr.woop))).all()
    sah_g, sah_r = orc.compact_sah(nodes, woop)["sah"], orc.compact_sah(ref.nodes, ref.woop)["sah"]
    assert abs(sah_g - sah_r) <= 0.005 * sah_r                   # north_star: builder SAH within 0.5 % (here: the same tree)
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    _check_against_oracle(orc, tracer, _frame_batches(gpu_host, tracer, scene, cam), nodes, woop, idx)


@pytest.mark.parametrize("kernel", ["b200_persistent_speculative_while_while", "b200_wide4"])
def test_bench_tree_at_full_size(gpu_host, orc, conference, kernel):
    verts, tris, cam, scene = conference
    capi.bvh_set_collapse(1, 8)
    try:
        bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(True, 2, 8, 0.001))       # bench.py's default tree
    finally:
        capi.bvh_set_collapse(0, 0)
    nodes, woop, idx = bvh.getNodeBuffer(), bvh.getTriWoopBuffer(), bvh.getTriIndexBuffer()
    c = orc.canonical(nodes, woop, idx)
    assert np.array_equal(np.sort(c.tris), np.arange(len(tris)))  # every triangle in exactly one leaf
    assert c.leaf_sizes.max() <= 8
    tracer = gpu_host.CudaBVHTracer()
    tracer.setKernel(kernel)
    tracer.setBVH(bvh)
    try:
        _check_against_oracle(orc, tracer, _frame_batches(gpu_host, tracer, scene, cam), nodes, woop, idx)
    finally:
        tracer.setKernel("b200_persistent_speculative_while_while")
