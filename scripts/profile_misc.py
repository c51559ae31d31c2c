"""Dev tool for ncu: one pass over every non-trace kernel family on the bench scene -- ray generation (primary, AO, shadow), hit count,
the ray Morton sort, the AOS/SOA -> Compact layout conversion, the HLBVH(2) + SAH-collapse build (the bench's tree), the device SAH
metric and the device Wide4 conversion -- one pass (`ncu --set full -k regex:<names> ...`).
Uses the checker only to produce an AOS_AOS BVH to upload (the reference's basic layout comes from its CPU builder)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402
import oracle  # noqa: E402  (checker: only to create the AOS_AOS input)


def main():
    import torch
    host.init(0)
    verts, tris, cam_name = scenes.config_scene("conference")
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    sv, st = scenes.room(60_000, seed=9, wall_frac=0.3)
    aos = oracle.CpuBVH(sv, st, oracle.BUILDER_SPLIT, 1, 8).basic(0)
    W, H = 1024, 768
    for rep in range(1):            # one pass: ncu replays every launch anyway
        capi.bvh_set_collapse(1, 8)
        capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, 2, 8, 0.001)
        capi.bvh_sah()
        capi.bvh_wide4_download()
        bvh = host.CudaBVH(layout=host.BVHLayout_Compact)
        bvh.resident = True
        tracer = host.CudaBVHTracer()
        tracer.setBVH(bvh)
        rg = host.RayGen(1 << 20)
        prim = host.RayBuffer()
        rg.primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
        tracer.traceBatch(prim)
        capi.count_hits(prim.getResultBuffer(), prim.getSize())
        rb = host.RayBuffer()
        rg.ao(rb, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED)
        rg.m_aoStartIdx = 0
        sh = host.RayBuffer()
        rg.shadow(sh, prim, 32, np.array([1.0, 2.0, 1.0], np.float32), 0.25, True, host.FIXED_AO_SEED)
        rg.m_shadowStartIdx = 0
        rb.mortonSort()
        # basic layout upload: rewritten on the device into the Compact form (nt_layout.cu)
        t2 = host.CudaBVHTracer()
        t2.setKernel("b200_persistent_speculative_while_while_aos_aos")
        t2.setBVH(host.CudaBVH(*aos, layout=0))
        t2.setKernel("b200_persistent_speculative_while_while")
        torch.cuda.synchronize()
    print("profile_misc done")


if __name__ == "__main__":
    main()
