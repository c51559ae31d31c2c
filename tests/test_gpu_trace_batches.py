"""GPU: nt_trace_batches (several batches in one persistent launch) gives exactly the results of one nt_trace_batch per batch."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(gpu_host):
    import torch
    verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
    scene = gpu_host.Scene(verts, tris)
    bvh = gpu_host.HLBVHBuilder(scene, gpu_host.HLBVHParams(True, 4, 8, 0.001))
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    cam = camera.named_camera("conference")
    prim = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 200, 120), 200, 120, cam.far)
    tracer.traceBatch(prim)
    rg = gpu_host.RayGen(1 << 14)                       # small batches: many of them
    bufs, new = [], True
    while True:
        rb = gpu_host.RayBuffer()
        ok, new = rg.ao(rb, prim, scene, 8, cam.far, new, gpu_host.FIXED_AO_SEED)
        if not ok:
            break
        bufs.append(rb.getRayBuffer().clone())
    assert len(bufs) > 8
    bufs[3] = bufs[3][:1000].clone()                    # ragged sizes
    bufs[5] = bufs[5][:1].clone()
    return gpu_host, tracer, bvh, bufs, torch


@pytest.mark.parametrize("kernel", ["b200_persistent_speculative_while_while", "b200_wide4", "b200_auto", "b200_persistent_speculative_while_while_fastmath"])
@pytest.mark.parametrize("closest", [True, False])
def test_one_launch_over_many_batches_equals_one_launch_per_batch(setup, kernel, closest):
    gpu_host, tracer, bvh, bufs, torch = setup
    capi.set_kernel(kernel)
    try:
        want = []
        for b in bufs:
            r = torch.full((len(b), 4), -9, dtype=torch.int32, device="cuda")
            capi.trace_batch(b, r, len(b), closest)
            want.append(r)
        got = [torch.full((len(b), 4), -9, dtype=torch.int32, device="cuda") for b in bufs]
        sec = capi.trace_batches(bufs, got, [len(b) for b in bufs], closest)
        assert sec > 0.0
        for w, g_ in zip(want, got):
            assert torch.equal(w, g_)
        # an empty batch in the list is skipped; more than 64 batches are split into several launches
        many = (bufs * 9)[:70]
        got2 = [torch.full((len(b), 4), -9, dtype=torch.int32, device="cuda") for b in many]
        counts = [len(b) for b in many]
        counts[2] = 0
        capi.trace_batches(many, got2, counts, closest)
        for i, (b, g_) in enumerate(zip(many, got2)):
            if i == 2:
                assert bool((g_ == -9).all())
            else:
                assert torch.equal(g_, want[i % len(bufs)])
        # queued form (nt_set_deferred): same results after nt_synchronize
        capi.set_deferred(2)
        got3 = [torch.full((len(b), 4), -9, dtype=torch.int32, device="cuda") for b in bufs]
        assert capi.trace_batches(bufs, got3, [len(b) for b in bufs], closest) == 0.0
        capi.synchronize()
        capi.set_deferred(0)
        for w, g_ in zip(want, got3):
            assert torch.equal(w, g_)
    finally:
        capi.set_deferred(0)
        capi.set_kernel("b200_persistent_speculative_while_while")


def test_trace_batches_refuses_host_buffers_and_other_kernels(setup):
    gpu_host, tracer, bvh, bufs, torch = setup
    res = torch.zeros((len(bufs[0]), 4), dtype=torch.int32, device="cuda")
    host_rays = bufs[0].cpu().numpy()
    with pytest.raises(capi.NtError, match="device buffers"):
        capi.trace_batches([host_rays], [res], [len(host_rays)], True)
    capi.set_kernel("b200_mr")
    try:
        with pytest.raises(capi.NtError, match="persistent one-ray kernel"):
            capi.trace_batches([bufs[0]], [res], [len(bufs[0])], True)
    finally:
        capi.set_kernel("b200_persistent_speculative_while_while")
