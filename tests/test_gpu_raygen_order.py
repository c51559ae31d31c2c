"""GPU: nt_raygen_set_order(1) writes the SAME rays as the reference order, permuted inside tiles of <= 1024 slots, with consistent
id <-> slot maps; tracing them gives the same result per ray id."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spp", [32, 7, 1])
def test_coherent_order_is_a_tile_local_permutation_with_identical_results(gpu_host, spp):
    verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
    scene = gpu_host.Scene(verts, tris)
    bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(True, 4, 8, 0.001))
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    cam = camera.named_camera("conference")
    w, h = 160, 96
    prim = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    tracer.traceBatch(prim)
    out = {}
    try:
        for order in (0, 1):
            capi.raygen_set_order(order)
            rg = gpu_host.RayGen(1 << 20)
            rb = gpu_host.RayBuffer()
            ok, _ = rg.ao(rb, prim, scene, spp, cam.far, True, gpu_host.FIXED_AO_SEED)
            assert ok
            rb.setNeedClosestHit(True)
            tracer.traceBatch(rb)
            out[order] = (rb.rays_host().copy(), rb.results_host().copy(), rb.getIDToSlotBuffer().cpu().numpy().copy(), rb.getSlotToIDBuffer().cpu().numpy().copy())
    finally:
        capi.raygen_set_order(0)
    rays0, res0, i2s0, s2i0 = out[0]
    rays1, res1, i2s1, s2i1 = out[1]
    n = len(rays0)
    assert np.array_equal(i2s0, np.arange(n)) and np.array_equal(s2i0, np.arange(n))           # the reference order is the identity
    assert np.array_equal(np.sort(i2s1), np.arange(n)) and np.array_equal(s2i1[i2s1], np.arange(n))   # a permutation and its inverse
    assert np.array_equal(rays1[i2s1], rays0)                                                  # ray id -> the same 32 bytes
    per_tile = (1024 // spp) * spp
    assert np.array_equal(i2s1 // per_tile, np.arange(n) // per_tile)                          # nothing leaves its tile
    assert not np.array_equal(i2s1, np.arange(n))                                              # and something did move
    assert np.array_equal(res1[i2s1], res0)                                                    # same hit, t, u, v per ray id
