"""BASELINE.json config 4: synthetic triangle soup, per-frame rebuild (vertices jittered every frame) + trace, swept over
leafSize {1,2,4,8} x builder {LBVH, HLBVH bits 4}.  Reports build ms / Mtris/s, primary and diffuse Mrays/s per cell and the
frame time (build + primary + one diffuse batch), plus size-independent checks per frame (every triangle once, sorted keys).
Usage: python scripts/rebuild_sweep.py [numTris=50000000] [frames=3]  -> gpurun_out/rebuild_sweep_<n>.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    host.init(0)
    verts, tris = scenes.soup_uniform(n, 5)
    lo, hi = scenes.bbox(verts)
    edge = float(n) ** (-1.0 / 3.0)
    d_verts0 = torch.from_numpy(verts).cuda()
    d_tris = torch.from_numpy(tris).cuda()
    cam = camera.look_at((1.6, 1.3, 1.1), (0.5, 0.5, 0.5), fov=60.0, near=0.001, far=10.0)
    prim = host.RayBuffer()
    host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
    sec_rays = host.RayBuffer()
    normals = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    out = {"num_tris": n, "frames": frames, "cells": []}
    for builder, bits, bname in ((capi.BUILDER_LBVH, 10, "LBVH"), (capi.BUILDER_HLBVH, 4, "HLBVH4")):
        for leaf in (1, 2, 4, 8):
            cell = {"builder": bname, "leafSize": leaf, "layout": "Compact", "build_ms": [], "primary_mrays": [], "diffuse_mrays": [], "frame_ms": []}
            for f in range(frames):
                g = torch.Generator(device="cuda"); g.manual_seed(1000 + f)
                d_verts = d_verts0 + (torch.rand(d_verts0.shape, generator=g, device="cuda") - 0.5) * (0.2 * edge)   # +-0.1 edge jitter
                torch.cuda.synchronize()
                flo = (d_verts.min(0).values.cpu().numpy()).astype(np.float32); fhi = (d_verts.max(0).values.cpu().numpy()).astype(np.float32)
                try:
                    b = capi.bvh_build(builder, d_verts, d_tris, flo, fhi, bits, leaf, 0.001)
                except capi.NtError as e:                              # BVHLayout_Compact child links are 32-bit BYTE offsets: < 1.98 GB of nodes
                    if "Compact2" not in str(e) or cell.get("layout") == "Compact2":
                        raise
                    capi.bvh_set_build_layout(5)                       # Compact2: offsets / 16
                    cell["layout"] = "Compact2"
                    b = capi.bvh_build(builder, d_verts, d_tris, flo, fhi, bits, leaf, 0.001)
                layout = 5 if cell.get("layout") == "Compact2" else 4
                bvh = host.CudaBVH(layout=layout); bvh.resident = True
                tracer = host.CudaBVHTracer()
                tracer.setKernel("b200_persistent_speculative_while_while_compact2" if layout == 5 else "b200_persistent_speculative_while_while")
                tracer.setBVH(bvh)
                tp = tracer.traceBatch(prim)
                capi.tri_normals(d_verts, d_tris, normals)
                hits = capi.count_hits(prim.getResultBuffer(), prim.getSize())
                n_in = (1 << 20) // 32
                first = prim.getSize() // 2 - n_in // 2                # slots from the middle of the image (the first slots see only sky)
                sec_rays.resize(n_in * 32)
                capi.raygen_ao(sec_rays.getRayBuffer(), sec_rays.getIDToSlotBuffer(), sec_rays.getSlotToIDBuffer(), prim.getRayBuffer(), prim.getResultBuffer(),
                               normals, first, n_in, 32, cam.far, 0x9E3779B9)
                sec_rays.setNeedClosestHit(True)
                td = tracer.traceBatch(sec_rays)
                hit_in = capi.count_hits(prim.getResultBuffer()[first:first + n_in], n_in)
                cell["build_ms"].append(b * 1e3); cell["primary_mrays"].append(prim.getSize() / tp * 1e-6)
                cell["diffuse_mrays"].append(hit_in * 32 / td * 1e-6); cell["frame_ms"].append((b + tp + td) * 1e3)
                if f == 0:                                             # size-independent checks, once per cell
                    keys, order = capi.bvh_build_debug(n)
                    assert (np.diff(keys.astype(np.int64)) >= 0).all(), "keys not sorted"
                    cnt = np.bincount(order, minlength=n)
                    assert cnt.min() == 1 and cnt.max() == 1, "every triangle exactly once"
                    cell["hit_fraction"] = hits / prim.getSize()
                    (nb, wb, ib), _ = capi.bvh_sizes()
                    cell["nodes"] = nb // 64; cell["bvh_mb"] = (nb + wb + ib) / 1e6
                    del keys, order, cnt
            capi.bvh_set_build_layout(4)
            for k in ("build_ms", "primary_mrays", "diffuse_mrays", "frame_ms"):
                cell[k] = float(np.mean(cell[k][1:] if frames > 1 else cell[k]))
            cell["build_mtris"] = n / cell["build_ms"] * 1e-3
            out["cells"].append(cell)
            print(json.dumps(cell), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/rebuild_sweep_{n}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
