"""GPU build time: the B200 builder vs the reference's own builder kernels recompiled for sm_100a (oracle/ref_gpu.py; its
-use_fast_math build, the reference's accounting: sum of kernel + thrust times), same scene, same GPU.
Usage: python scripts/ref_build_compare.py -> gpurun_out/ref_build_compare.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_gpu  # noqa: E402
import oracle  # noqa: E402
from ntrace_b200 import capi, host, scenes  # noqa: E402


def main():
    host.init(0)
    rows = []
    for name, gen in (("conference 283K", lambda: scenes.config_scene("conference")[:2]), ("fairyforest 174K", lambda: scenes.config_scene("fairyforest")[:2]),
                      ("soup 2M", lambda: scenes.soup_uniform(2_000_000, 5)), ("room 10.5M", lambda: scenes.config_scene("sanmiguel")[:2])):
        verts, tris = gen()
        lo, hi = scenes.bbox(verts)
        scene = host.Scene(verts, tris)
        row = {"scene": name, "num_tris": int(len(tris))}
        for label, builder, bits in (("lbvh", capi.BUILDER_LBVH, 10), ("hlbvh4", capi.BUILDER_HLBVH, 4)):
            mine = min(capi.bvh_build(builder, scene.vtxPos, scene.triVtxIndex, lo, hi, bits, 8, 0.001) for _ in range(4)) * 1e3
            nodes, woop, _, _ = capi.bvh_download()
            fn = ref_gpu.lbvh_build if label == "lbvh" else (lambda *a, **k: ref_gpu.hlbvh_build(a[0], a[1], a[2], a[3], 4, *a[4:], **k))
            ref_ms, ref_sah = [], None
            for _ in range(3):
                r = fn(scene.vtxPos, scene.triVtxIndex, lo, hi, 8, 0.001, ieee=False)
                ref_ms.append(r["gpu_ms"])
            ref_sah = oracle.compact_sah(r["nodes"], r["woop"])["sah"]
            row[label] = {"b200_ms": mine, "reference_kernels_ms": min(ref_ms), "speedup": min(ref_ms) / mine,
                          "b200_sah": oracle.compact_sah(nodes, woop)["sah"], "reference_sah": ref_sah}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del scene
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/ref_build_compare.json", "w"), indent=1)


if __name__ == "__main__":
    main()
