"""Full-size configs of BASELINE.json (San Miguel stand-in 10.5M tris, 50M-triangle soup): GPU build timings and
size-independent correctness properties (every triangle exactly once, sorted keys, root box = scene box +- eps,
LBVH vs HLBVH hit agreement, full canonical parity against the CPU restatement where it finishes in seconds).
Usage: python scripts/large_scene_check.py [sanmiguel|soup50m|all]   (writes gpurun_out/large_<name>.json)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker)
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402


def check(name):
    import torch
    t0 = time.time()
    verts, tris, cam_name = scenes.config_scene(name)
    gen_s = time.time() - t0
    n = len(tris)
    lo, hi = scenes.bbox(verts)
    scene = host.Scene(verts, tris)
    out = {"scene": name, "num_tris": n, "gen_s": gen_s}
    cam = camera.named_camera(cam_name) if cam_name != "soup" else camera.look_at((1.6, 1.3, 1.1), (0.5, 0.5, 0.5), fov=60.0, near=0.001, far=10.0)
    rays = host.RayBuffer()
    host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
    tracer = host.CudaBVHTracer()
    results = {}
    for label, builder, bits in (("lbvh", capi.BUILDER_LBVH, 10), ("hlbvh", capi.BUILDER_HLBVH, 4)):
        ts = [capi.bvh_build(builder, scene.vtxPos, scene.triVtxIndex, lo, hi, bits, 8, 0.001) for _ in range(4)]
        (nb, wb, ib), _ = capi.bvh_sizes()
        out[label] = {"build_ms_best": min(ts[1:]) * 1e3, "build_ms_first": ts[0] * 1e3, "mtris_per_s": n / min(ts[1:]) * 1e-6,
                      "nodes": nb // 64, "bvh_mb": (nb + wb + ib) / 1e6}
        nodes, woop, idx, _ = capi.bvh_download()
        keys, order = capi.bvh_build_debug(n)
        assert (np.diff(keys.astype(np.int64)) >= 0).all(), "keys not sorted"
        assert np.array_equal(np.sort(order), np.arange(n, dtype=np.int32)), "sort lost triangles"
        w = woop.reshape(-1, 4)
        term = (w == np.int32(-2147483648)).all(1)
        # layout invariant: woop has 3n triangle rows + numLeaves terminators, triIndex holds [id,0,0] per triangle
        num_leaves = int(term.sum())
        assert len(w) == 3 * n + num_leaves
        nonterm_idx = idx[~term].reshape(-1, 3)
        assert (nonterm_idx[:, 1:] == 0).all()
        assert np.array_equal(np.sort(nonterm_idx[:, 0]), np.arange(n, dtype=np.int32)), "every triangle exactly once"
        f = nodes.view(np.float32).reshape(-1, 16)
        rlo = np.array([min(f[0, 0], f[0, 4]), min(f[0, 2], f[0, 6]), min(f[0, 8], f[0, 10])])
        rhi = np.array([max(f[0, 1], f[0, 5]), max(f[0, 3], f[0, 7]), max(f[0, 9], f[0, 11])])
        used_lo, used_hi = verts[np.unique(tris)].min(0), verts[np.unique(tris)].max(0)
        assert np.allclose(rlo, used_lo - 0.001, atol=1e-5) and np.allclose(rhi, used_hi + 0.001, atol=1e-5), "root box"
        bvh = host.CudaBVH(layout=host.BVHLayout_Compact); bvh.resident = True
        tracer.setBVH(bvh)
        for _ in range(2):
            tracer.traceBatch(rays)
        sec = np.mean([tracer.traceBatch(rays) for _ in range(5)])
        out[label]["primary_mrays"] = rays.getSize() / sec * 1e-6
        results[label] = rays.results_host().copy()
        if n <= 12_000_000 and label == "lbvh":
            t0 = time.time()
            ref = oracle.lbvh_build(verts, tris, lo, hi, hlbvh=False, leaf_size=8)
            cg, cr = oracle.canonical(nodes, woop, idx), oracle.canonical(ref.nodes, ref.woop, ref.tri_index)
            out[label]["oracle_parity"] = bool(np.array_equal(keys, ref.sorted_keys) and np.array_equal(order, ref.sorted_idx)
                                               and np.array_equal(cg.inner, cr.inner) and np.array_equal(cg.boxes, cr.boxes)
                                               and np.array_equal(cg.tris, cr.tris))
            out[label]["oracle_s"] = time.time() - t0
            assert out[label]["oracle_parity"]
        del nodes, woop, idx, w
    a, b = results["lbvh"], results["hlbvh"]
    same = a[:, 0] == b[:, 0]
    ta, tb = a[:, 1].view(np.float32), b[:, 1].view(np.float32)
    rel = np.abs(ta - tb) / np.maximum(np.abs(tb), 1e-30)
    out["lbvh_vs_hlbvh_id_match"] = float(same.mean())
    out["lbvh_vs_hlbvh_nontie_mismatch"] = float(((~same) & (rel > 1e-4)).mean())
    assert out["lbvh_vs_hlbvh_nontie_mismatch"] <= 1e-4
    out["hit_fraction"] = float((a[:, 0] >= 0).mean())
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/large_{name}.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    host.init(0)
    for nm in (["sanmiguel", "soup50m"] if which == "all" else [which]):
        check(nm)
