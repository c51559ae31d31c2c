"""Dev tool (not the contract bench): Mrays/s of several trace kernels on the bench frame (Conference stand-in, GPU HLBVH(2) + collapse),
per ray type, one synchronous launch at a time (CUDA events around each launch), plus parity of each kernel against the first one and,
with --check, against the oracle on a strided sample.  Usage:
    python scripts/kernel_compare.py --kernels b200_persistent_speculative_while_while,b200_wide4 [--batches 6] [--check] [--out file.json]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402

W, H, SPP = 1024, 768, 32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernels", default="b200_persistent_speculative_while_while,b200_wide4")
    ap.add_argument("--scene", default="conference")
    ap.add_argument("--batches", type=int, default=24, help="secondary batches per ray type that are timed")
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--hlbvh-bits", type=int, default=2)
    ap.add_argument("--collapse", type=int, default=1)
    ap.add_argument("--max-leaf", type=int, default=8)
    ap.add_argument("--leaf", type=int, default=8)
    ap.add_argument("--raygen-order", type=int, default=0, help="nt_raygen_set_order: 0 reference slot order, 1 direction-coherent tiles")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    host.init(0)
    verts, tris, cam_name = scenes.config_scene(args.scene)
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    capi.bvh_set_collapse(args.collapse, args.max_leaf)
    build_s = capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, args.hlbvh_bits, args.leaf, 0.001)
    bvh = host.CudaBVH(layout=host.BVHLayout_Compact)
    bvh.resident = True
    tracer = host.CudaBVHTracer()
    tracer.setBVH(bvh)
    capi.raygen_set_order(args.raygen_order)
    rg = host.RayGen(1 << 20)
    prim = host.RayBuffer()
    rg.primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
    tracer.traceBatch(prim)
    hits = capi.count_hits(prim.getResultBuffer(), prim.getSize())
    batches = [("primary", prim.getRayBuffer().clone(), prim.getSize(), True)]
    for name, dist_max, closest in (("AO", 5.0, False), ("diffuse", cam.far, True)):
        new, k = True, 0
        while k < args.batches:
            rb = host.RayBuffer()
            ok, new = rg.ao(rb, prim, scene, SPP, dist_max, new, host.FIXED_AO_SEED)
            if not ok:
                break
            batches.append((name, rb.getRayBuffer(), rb.getSize(), closest))
            k += 1
        rg.m_aoStartIdx = 0
    out = {"scene": args.scene, "tris": int(len(tris)), "build_ms": build_s * 1e3, "primary_hits": int(hits), "kernels": {}}
    first = None
    kernels = args.kernels.split(",")
    res = {k: [torch.zeros((n, 4), dtype=torch.int32, device="cuda") for _, _, n, _ in batches] for k in kernels}
    for kname in kernels:
        capi.set_kernel(kname)
        sec = {"primary": 0.0, "AO": 0.0, "diffuse": 0.0}
        rays_n = {"primary": 0, "AO": 0, "diffuse": 0}
        for i, (name, rays, n, closest) in enumerate(batches):
            capi.trace_batch(rays, res[kname][i], n, closest)          # warm-up (also derives the wide form)
            ts = [capi.trace_batch(rays, res[kname][i], n, closest) for _ in range(args.repeats)]
            sec[name] += float(np.mean(ts))
            rays_n[name] += n
        row = {t: rays_n[t] / sec[t] * 1e-6 for t in sec if sec[t] > 0}
        # parity against the first kernel of the list
        if first is None:
            first = kname
        else:
            par = {}
            for t, closest in (("primary", True), ("AO", False), ("diffuse", True)):
                a = torch.cat([res[first][i] for i, b in enumerate(batches) if b[0] == t])
                b = torch.cat([res[kname][i] for i, bb in enumerate(batches) if bb[0] == t])
                ha, hb = a[:, 0] >= 0, b[:, 0] >= 0
                d = {"flag_match": float((ha == hb).float().mean()), "id_match": float((a[:, 0] == b[:, 0]).float().mean())}
                if closest:
                    ta, tb = a[:, 1].view(torch.float32), b[:, 1].view(torch.float32)
                    both = ha & hb
                    d["t_bit_equal_on_hits"] = float((a[:, 1][both] == b[:, 1][both]).float().mean()) if both.any() else 1.0
                    rel = ((ta - tb).abs() / tb.abs().clamp_min(1e-30))[both]
                    d["max_rel_t"] = float(rel.max()) if both.any() else 0.0
                    mism = (a[:, 0] != b[:, 0]) & both
                    d["id_mismatch_max_rel_t"] = float(((ta - tb).abs() / tb.abs().clamp_min(1e-30))[mism].max()) if mism.any() else 0.0
                par[t] = d
            row["parity_vs_" + first] = par
        out["kernels"][kname] = row
        print(kname, json.dumps(row), flush=True)
    if args.check:
        import oracle  # checker only
        nodes, woop, idx, layout = capi.bvh_download()
        wn, depth = capi.bvh_wide4_convert_host(layout, nodes, woop.nbytes)
        chk = {}
        for kname in kernels:
            chk[kname] = {}
            for t, closest in (("primary", True), ("AO", False), ("diffuse", True)):
                i = [j for j, b in enumerate(batches) if b[0] == t][0]
                stride = max(1, batches[i][2] // 200_000)
                rays = batches[i][1][::stride].cpu().numpy()
                got = res[kname][i][::stride].cpu().numpy()
                ref = oracle.wide4_trace(wn, woop, idx, rays, closest) if "wide4" in kname else oracle.compact_trace(nodes, woop, idx, rays, closest)
                chk[kname][t] = {"rays": int(len(rays)), "id_equal": float((got[:, 0] == ref[:, 0]).mean()), "t_bits_equal": float((got[:, 1] == ref[:, 1]).mean())}
        out["oracle_check"] = chk
        out["wide_depth"] = depth
        out["wide_nodes"] = int(len(wn) // 16)
        out["binary_nodes"] = int(len(nodes) // 16)
        print("oracle_check", json.dumps(chk), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
