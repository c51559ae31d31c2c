"""Launch sequence for ncu captures of ONE trace kernel on the bench frame (dev tool; run under ncu by scripts/ncu_binding.py).
After setup it issues exactly four launches of the chosen kernel through nt_trace_batch, in this order:
    primary (warm-up, also produces the primary hits), primary, AO batch `--batch`, diffuse batch `--batch`
so `ncu -k regex:trace_ --launch-skip <setup launches of that name>` sees them in a fixed order.  Usage:
    python scripts/profile_kernel.py --kernel b200_wide4 [--scene conference] [--batch 12]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="b200_persistent_speculative_while_while")
    ap.add_argument("--scene", default="conference")
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--hlbvh-bits", type=int, default=2)
    ap.add_argument("--collapse", type=int, default=1)
    ap.add_argument("--max-leaf", type=int, default=8)
    ap.add_argument("--raygen-order", type=int, default=1)
    args = ap.parse_args()
    import torch
    host.init(0)
    verts, tris, cam_name = scenes.config_scene(args.scene)
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    capi.bvh_set_collapse(args.collapse, args.max_leaf)
    capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, args.hlbvh_bits, 8, 0.001)
    bvh = host.CudaBVH(layout=host.BVHLayout_Compact)
    bvh.resident = True
    tracer = host.CudaBVHTracer()
    tracer.setKernel(args.kernel)
    tracer.setBVH(bvh)
    W, H = 1024, 768
    capi.raygen_set_order(args.raygen_order)
    rg = host.RayGen(1 << 20)
    prim = host.RayBuffer()
    rg.primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
    tracer.traceBatch(prim)                                       # launch 1 (warm-up; primary hits feed the secondary batches)
    sec = {}
    for name, dist_max in (("AO", 5.0), ("diffuse", cam.far)):
        new = True
        for k in range(args.batch + 1):
            rb = host.RayBuffer()
            ok, new = rg.ao(rb, prim, scene, 32, dist_max, new, host.FIXED_AO_SEED)
            if not ok:
                break
            sec[name] = rb
        rg.m_aoStartIdx = 0
    res = torch.zeros((1 << 20, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = [capi.trace_batch(prim.getRayBuffer(), res, prim.getSize(), True),                     # launch 2
         capi.trace_batch(sec["AO"].getRayBuffer(), res, sec["AO"].getSize(), False),           # launch 3
         capi.trace_batch(sec["diffuse"].getRayBuffer(), res, sec["diffuse"].getSize(), True)]  # launch 4
    print("kernel", args.kernel, "ms", [x * 1e3 for x in t], "rays", [prim.getSize(), sec["AO"].getSize(), sec["diffuse"].getSize()], flush=True)


if __name__ == "__main__":
    main()
