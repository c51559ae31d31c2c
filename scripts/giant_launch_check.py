"""dev tool: how much of a frame's time is launch ramp-up / drain that the two-stream overlap does not already hide?  The 24 diffuse (and 24 AO)
batches of the bench frame traced (a) one launch per batch, synchronous, (b) queued on two kernel streams (nt_set_deferred(2)), (c) as ONE
launch over the concatenated buffer."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402

import torch
host.init(0)
verts, tris, cam_name = scenes.config_scene("conference")
cam = camera.named_camera(cam_name)
scene = host.Scene(verts, tris)
capi.bvh_set_collapse(1, 8)
capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, 2, 8, 0.001)
bvh = host.CudaBVH(layout=host.BVHLayout_Compact); bvh.resident = True
tracer = host.CudaBVHTracer(); tracer.setKernel("b200_auto"); tracer.setBVH(bvh)
capi.raygen_set_order(1)
rg = host.RayGen(1 << 20)
prim = host.RayBuffer()
rg.primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far, 0)
tracer.traceBatch(prim)
for name, dist_max, closest in (("AO", 5.0, False), ("diffuse", cam.far, True)):
    bufs, new = [], True
    while True:
        rb = host.RayBuffer()
        ok, new = rg.ao(rb, prim, scene, 32, dist_max, new, host.FIXED_AO_SEED)
        if not ok:
            break
        bufs.append(rb.getRayBuffer().clone())
    rg.m_aoStartIdx = 0
    n = sum(len(b) for b in bufs)
    big = torch.cat(bufs)
    res = [torch.zeros((len(b), 4), dtype=torch.int32, device="cuda") for b in bufs]
    res_big = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        a = sum(capi.trace_batch(b, r, len(b), closest) for b, r in zip(bufs, res))
    capi.set_deferred(2)
    for rep in range(2):
        capi.event_record(0)
        for b, r in zip(bufs, res):
            capi.trace_batch(b, r, len(b), closest)
        capi.event_record(1)
        bsec = capi.event_elapsed(0, 1)
    capi.set_deferred(0)
    for _ in range(2):
        c = capi.trace_batch(big, res_big, n, closest)
    same = bool(torch.equal(torch.cat(res), res_big))
    print(f"{name}: {len(bufs)} launches synchronous {n / a * 1e-6:.0f} Mrays/s, two kernel streams {n / bsec * 1e-6:.0f}, one launch of {n} rays {n / c * 1e-6:.0f}; results equal: {same}", flush=True)
