"""BASELINE.json configs 0-2 (Sibenik / Conference / Fairy Forest stand-ins) through the reference's Renderer loop on one
B200: GPU build, Mrays/s per ray type unsorted and sorted (kernel time only, the reference's accounting), and parity of a
ray sample of every type against the oracle (flat Woop tracer on the same GPU-built buffers).
Usage: python scripts/config_matrix.py [names...]   -> gpurun_out/config_matrix.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker)
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402

W, H, SPP = 1024, 768, 32


def run(name):
    verts, tris, cam_name = scenes.config_scene(name)
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    out = {"scene": name, "num_tris": int(len(tris)), "camera": cam_name}
    for label, builder, bits, collapse in (("lbvh", "LBVH", 10, 0), ("hlbvh4", "HLBVH", 4, 0), ("hlbvh2_collapse", "HLBVH", 2, 1)):
        capi.bvh_set_collapse(collapse, 8 if collapse else 0)
        r = host.Renderer(host.BuildSettings(builder=builder, hlbvh=host.HLBVHParams(True, bits, 8, 0.001)))
        r.setScene(scene)
        row = {}
        bvh = None
        for rt in (host.RayType_Primary, host.RayType_AO, host.RayType_Diffuse):
            for sort in (False, True):
                if rt == host.RayType_Primary and sort:
                    continue
                aor = 5.0 if name != "fairyforest" else 0.05 * float(np.linalg.norm(verts.max(0) - verts.min(0)))
                r.setParams(host.RendererParams(rayType=rt, numSamples=SPP, aoRadius=aor, sortSecondary=sort))
                r.beginFrame(cam, W, H)
                if bvh is None:
                    bvh = r.getCudaBVH()
                    lo, hi = scene.getBBox()                               # steady-state build time: best of 3 rebuilds (the first call also grows scratch)
                    bk = capi.BUILDER_LBVH if builder == "LBVH" else capi.BUILDER_HLBVH
                    row["build_ms"] = min(capi.bvh_build(bk, scene.vtxPos, scene.triVtxIndex, lo, hi, bits, 8, 0.001) for _ in range(3)) * 1e3
                    row["build_mtris"] = len(tris) / row["build_ms"] * 1e-3
                    nodes, woop, idx, _ = capi.bvh_download()
                    row["sah"] = oracle.compact_sah(nodes, woop)["sah"]
                counted = r.getTotalNumRays()
                sec, sample = 0.0, None
                while r.nextBatch():
                    if sort and "sort_ms_per_mi_rays" not in row:              # cost of mortonSort itself, per 2^20 rays
                        import torch
                        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize(); t0.record()
                        capi.synchronize(); r.m_batchRays.mortonSort(); capi.synchronize()
                        t1.record(); torch.cuda.synchronize()
                        row["sort_ms_per_mi_rays"] = t0.elapsed_time(t1) / r.m_batchRays.getSize() * (1 << 20)
                    r.traceBatch()
                    ts = [r.traceBatch() for _ in range(3)]
                    sec += float(np.mean(ts))
                    if sample is None:
                        sample = (r.m_batchRays.rays_host()[:200_000].copy(), r.m_batchRays.results_host()[:200_000].copy(), r.m_batchRays.getNeedClosestHit())
                key = rt + ("_sorted" if sort else "")
                row[key + "_mrays"] = counted / sec * 1e-6
                if not sort:                                                 # parity of the first 200K rays of the first batch
                    rays, got, closest = sample
                    want = oracle.compact_trace(nodes, woop, idx, rays, closest)
                    live = rays[:, 7] >= rays[:, 3]
                    if closest:
                        same = got[live, 0] == want[live, 0]
                        hit = same & (want[live, 0] >= 0)
                        tg, tw = got[live, 1].view(np.float32)[hit], want[live, 1].view(np.float32)[hit]
                        row[key + "_id_match"] = float(same.mean())
                        row[key + "_max_rel_t"] = float(np.max(np.abs(tg - tw) / np.maximum(np.abs(tw), 1e-30))) if hit.any() else 0.0
                        assert same.mean() >= 0.9999 and row[key + "_max_rel_t"] <= 1e-5
                    else:
                        row[key + "_hit_match"] = float(((got[live, 0] >= 0) == (want[live, 0] >= 0)).mean())
                        assert row[key + "_hit_match"] >= 0.9999
        out[label] = row
    capi.bvh_set_collapse(0, 0)
    # CPU reference leg of config 0: restated SplitBVH + BVH::trace (what the config names), bounded to this scene size
    if name == "sibenik":
        import time
        t0 = time.time()
        cpu = oracle.CpuBVH(verts, tris, oracle.BUILDER_SPLIT, 1, 1, 1.0e-5)
        out["cpu_splitbvh_build_s"] = time.time() - t0
        out["cpu_splitbvh_sah"] = cpu.stats().sah
        rays, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far)
        t0 = time.time()
        res = cpu.trace(rays, True, nthreads=oracle.max_threads())
        out["cpu_trace_mrays"] = len(rays) / (time.time() - t0) * 1e-6
        out["cpu_threads"] = oracle.max_threads()
        tracer = host.CudaBVHTracer()
        tracer.setBVH(host.CudaBVH(*cpu.compact()))
        rb = host.RayBuffer(); rb.setRays(rays)
        tracer.traceBatch(rb)
        sec = float(np.mean([tracer.traceBatch(rb) for _ in range(5)]))
        got = rb.results_host()
        out["gpu_on_cpu_splitbvh_mrays"] = len(rays) / sec * 1e-6
        out["gpu_vs_cpu_tree_trace_id_match"] = float((got[:, 0] == res[:, 0]).mean())
    return out


if __name__ == "__main__":
    host.init(0)
    names = sys.argv[1:] or ["sibenik", "conference", "fairyforest"]
    res = [run(n) for n in names]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/config_matrix.json", "w"), indent=1)
    for r in res:
        print(json.dumps(r))
