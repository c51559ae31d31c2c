"""B200 kernel vs the reference's own traversal kernels recompiled for sm_100a (oracle/ref_gpu.py): result agreement statistics
and Mrays/s of both on the same GPU, BVH and rays.  Usage: python scripts/ref_kernel_compare.py [scene=conference]
-> gpurun_out/ref_kernel_compare_<scene>.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_gpu  # noqa: E402  (checker / baseline)
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "conference"
    host.init(0)
    verts, tris, cam_name = scenes.config_scene(name)
    cam = camera.named_camera(cam_name)
    diag = float(np.linalg.norm(verts.max(0) - verts.min(0)))
    scene = host.Scene(verts, tris)
    lo, hi = scene.getBBox()
    out = {"scene": name, "num_tris": int(len(tris)), "rows": []}
    prim = host.RayBuffer()
    host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
    capi.bvh_set_collapse(1, 8)
    capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 2, 8, 0.001)       # the bench's BVH
    capi.bvh_set_collapse(0, 0)
    tracer = host.CudaBVHTracer()
    bvh = host.CudaBVH(layout=4); bvh.resident = True
    tracer.setBVH(bvh)
    tracer.traceBatch(prim)
    ao, diff = host.RayBuffer(), host.RayBuffer()
    host.RayGen(1 << 20).ao(ao, prim, scene, 32, 5.0, True, host.FIXED_AO_SEED)
    host.RayGen(1 << 20).ao(diff, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED)
    diff.setNeedClosestHit(True)
    batches = {"primary": prim, "AO": ao, "diffuse": diff}
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for kernel, layout in (("fermi_speculative_while_while", 4), ("kepler_dynamic_fetch", 5)):
        capi.bvh_convert(layout)
        nodes, woop, idx, _ = capi.bvh_download()
        d_nodes, d_woop, d_idx = (torch.from_numpy(a).cuda() for a in (nodes, woop, idx))
        tracer.setKernel(kernel)
        b = host.CudaBVH(layout=layout); b.resident = True
        tracer.setBVH(b)
        for rt, rb in batches.items():
            closest = rb.getNeedClosestHit()
            rays = rb.rays_host()
            live = rays[:, 7] >= rays[:, 3]
            tracer.traceBatch(rb)
            mine_s = min(tracer.traceBatch(rb) for _ in range(5))
            mine = rb.results_host().copy()
            res = torch.zeros((rb.getSize(), 4), dtype=torch.int32, device="cuda")
            row = {"kernel": kernel, "ray_type": rt, "rays": int(rb.getSize()), "b200_mrays": rb.getSize() / mine_s * 1e-6}
            # as shipped: one warp per 32 rays (fermi) or the hard-coded 720 persistent warps (kepler); then the persistent
            # launch re-sized for this GPU (kernel code untouched): 148 SMs x resident warps
            ms, cfg = ref_gpu.trace(kernel, rb.getRayBuffer(), res, d_nodes, d_woop, d_idx, any_hit=not closest, repeats=5)
            row["reference_as_shipped_mrays"] = rb.getSize() / ms * 1e-3
            row["reference_cfg"] = cfg
            if cfg["usePersistentThreads"]:
                best = 0.0
                for wps in (16, 32, 48, 64):
                    ms2, _ = ref_gpu.trace(kernel, rb.getRayBuffer(), res, d_nodes, d_woop, d_idx, any_hit=not closest, desired_warps=sms * wps, repeats=3)
                    best = max(best, rb.getSize() / ms2 * 1e-3)
                row["reference_resized_launch_mrays"] = best
            want = res.cpu().numpy()
            g, w = mine[live], want[live]
            if closest:
                same = g[:, 0] == w[:, 0]
                tg, tw = g[:, 1].view(np.float32), w[:, 1].view(np.float32)
                mm = ~same
                both = mm & (g[:, 0] >= 0) & (w[:, 0] >= 0)
                non_tie = np.zeros(len(g), bool)
                non_tie[both] = np.abs(tg[both] - tw[both]) > 1e-4 * np.maximum(np.abs(tw[both]), 1e-30)
                non_tie |= mm & ((g[:, 0] >= 0) != (w[:, 0] >= 0))
                hit = same & (w[:, 0] >= 0)
                err = np.abs(tg[hit] - tw[hit]); rel = err / np.maximum(np.abs(tw[hit]), 1e-30)
                row.update(id_match=float(same.mean()), non_tie_mismatch=int(non_tie.sum()), t_rel_median=float(np.median(rel)),
                           t_rel_q999=float(np.quantile(rel, 0.999)), t_rel_max=float(rel.max()), t_abs_max_over_diag=float(err.max() / diag),
                           frac_rel_gt_1e5=float((rel > 1e-5).mean()))
                k = int(np.argmax(rel))
                row["worst"] = {"t_b200": float(tg[hit][k]), "t_ref": float(tw[hit][k])}
            else:
                row.update(hit_match=float(((g[:, 0] >= 0) == (w[:, 0] >= 0)).mean()))
            out["rows"].append(row)
            print(json.dumps(row), flush=True)
    capi.bvh_convert(4)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/ref_kernel_compare_{name}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
