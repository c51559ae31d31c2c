"""nt_shutdown / nt_init cycle (runs last: it drops the library's device state).  The builder's and the ray sorter's grow-only
scratch must go with the context -- it belongs to the device it was allocated on -- and a second context must build and trace
exactly like the first."""
import numpy as np
import pytest

from ntrace_b200 import camera, capi, scenes

pytestmark = pytest.mark.gpu


def _frame(gpu_host, verts, tris):
    scene = gpu_host.Scene(verts, tris)
    bvh = gpu_host.HLBVHBuilder(scene)
    tracer = gpu_host.CudaBVHTracer()
    tracer.setBVH(bvh)
    cam = camera.named_camera("conference")
    rb = gpu_host.RayBuffer()
    gpu_host.RayGen().primary(rb, cam.position, camera.nscreen_to_world(cam, 128, 96), 128, 96, cam.far)
    tracer.traceBatch(rb)
    sec = gpu_host.RayBuffer()
    gpu_host.RayGen().ao(sec, rb, scene, 4, cam.far, True, gpu_host.FIXED_AO_SEED)
    sec.mortonSort()
    return rb.results_host().copy()


def test_shutdown_releases_scratch_and_a_new_context_repeats_the_results(gpu_host):
    import torch
    verts, tris = scenes.room(300_000, seed=5, wall_frac=0.3)
    first = _frame(gpu_host, verts, tris)
    torch.cuda.synchronize()
    free_with_scratch = torch.cuda.mem_get_info()[0]
    capi.shutdown()
    free_after = torch.cuda.mem_get_info()[0]
    # a 300 K-triangle HLBVH build holds > 40 MB of scratch (keys, indices, ranges, per-triangle boxes, bins)
    assert free_after - free_with_scratch > 40 << 20, "nt_shutdown left the builder / sorter scratch allocated"
    with pytest.raises(capi.NtError):
        capi.bvh_sizes()                                   # no context: every call fails, nothing falls back
    gpu_host.init(0)
    second = _frame(gpu_host, verts, tris)
    assert np.array_equal(first.view(np.uint32), second.view(np.uint32))
