"""BASELINE.json configs 0-2 (Sibenik / Conference / Fairy Forest stand-ins) with the round-2 defaults on one B200: bench tree
(HLBVH(2) + SAH collapse), kernel b200_auto, direction-coherent slot order, through the reference-shaped Renderer: Mrays/s per ray type
with one launch per <= 1 Mi-ray batch (the reference's loop) and with one launch per ray type (prepareFrame / traceFrame), kernel time only,
plus parity of a strided ray sample of every type against the oracle (flat Woop tracer on the GPU-built buffers).
Usage: python scripts/config_matrix_r2.py [out.json]"""
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
    capi.bvh_set_collapse(1, 8)
    capi.raygen_set_order(1)
    r = host.Renderer(host.BuildSettings(builder="HLBVH", hlbvh=host.HLBVHParams(True, 2, 8, 0.001)))
    r.setScene(scene)
    out = {"scene": name, "num_tris": int(len(tris)), "camera": cam_name}
    aor = 5.0 if name != "fairyforest" else 0.05 * float(np.linalg.norm(verts.max(0) - verts.min(0)))
    nodes = None
    for rt, label in ((host.RayType_Primary, "primary"), (host.RayType_AO, "AO"), (host.RayType_Diffuse, "diffuse")):
        r.setParams(host.RendererParams(kernelName="b200_auto", rayType=rt, numSamples=SPP, aoRadius=aor, sortSecondary=False))
        r.beginFrame(cam, W, H)
        if nodes is None:
            lo, hi = scene.getBBox()
            out["build_ms"] = min(capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 2, 8, 0.001) for _ in range(4)) * 1e3
            out["tree"] = capi.bvh_sah()
            r.m_bvh = None
            r.beginFrame(cam, W, H)
            nodes, woop, idx, _ = capi.bvh_download()
        counted = r.getTotalNumRays()
        sec = 0.0
        while r.nextBatch():
            r.traceBatch()
            sec += float(np.mean([r.traceBatch() for _ in range(3)]))
        row = {"counted_rays": int(counted), "mrays_per_batch_launch": counted / sec * 1e-6}
        r.beginFrame(cam, W, H)
        r.prepareFrame()
        r.traceFrame()
        row["mrays_frame_launch"] = counted / float(np.mean([r.traceFrame() for _ in range(3)])) * 1e-6
        rb = r.getFrameBatches()[len(r.getFrameBatches()) // 2]
        rays, got = rb.rays_host(), rb.results_host()
        stride = max(1, len(rays) // 200_000)
        rays, got = np.ascontiguousarray(rays[::stride]), got[::stride]
        closest = rt != host.RayType_AO
        ref = oracle.compact_trace(nodes, woop, idx, rays, closest)
        row["parity_rays"] = int(len(rays))
        row["flag_match"] = float(((got[:, 0] >= 0) == (ref[:, 0] >= 0)).mean())
        if closest:
            row["id_match"] = float((got[:, 0] == ref[:, 0]).mean())
            hit = (got[:, 0] >= 0) & (ref[:, 0] >= 0)
            tg, tr = got[:, 1].view(np.float32), ref[:, 1].view(np.float32)
            row["max_rel_t"] = float(np.max(np.abs(tg - tr)[hit] / np.abs(tr[hit]))) if hit.any() else 0.0
        out[label] = row
    capi.bvh_set_collapse(0, 0)
    capi.raygen_set_order(0)
    return out


if __name__ == "__main__":
    host.init(0)
    res = [run(n) for n in ("sibenik", "conference", "fairyforest")]
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
