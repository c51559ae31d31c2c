"""Early measurement helper (not the contract bench): Mrays/s of the trace kernels on the conference
stand-in with a CPU-built SplitBVH, primary / AO / diffuse.  Usage: python scripts/quick_trace_bench.py [ntris]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker-side BVH builder; this script is a dev tool, not the product path)
from ntrace_b200 import camera, host, scenes  # noqa: E402


def main():
    ntris = int(sys.argv[1]) if len(sys.argv) > 1 else 283_000
    host.init(0)
    verts, tris = scenes.room(ntris, seed=2, wall_frac=0.3)
    cam = camera.named_camera("conference")
    w, h = 1024, 768
    t0 = time.time()
    cpu = oracle.CpuBVH(verts, tris, oracle.BUILDER_SPLIT, 1, 1)
    st = cpu.stats()
    print(f"SplitBVH {ntris} tris: {time.time() - t0:.1f}s sah={st.sah:.2f} inner={st.num_inner} dup={st.duplicates}", flush=True)
    nodes, woop, idx = cpu.compact()
    scene = host.Scene(verts, tris)
    tracer = host.CudaBVHTracer()
    tracer.setBVH(host.CudaBVH(nodes, woop, idx))
    rg = host.RayGen()
    prim = host.RayBuffer()
    rg.primary(prim, cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
    for kernel in ["b200_persistent_speculative_while_while", "b200_speculative_while_while"]:
        tracer.setKernel(kernel)
        for _ in range(3):
            tracer.traceBatch(prim)
        ts = [tracer.traceBatch(prim) for _ in range(10)]
        print(f"{kernel}: primary {prim.getSize() / np.mean(ts) * 1e-6:.1f} Mrays/s (best {prim.getSize() / np.min(ts) * 1e-6:.1f})", flush=True)
        hits = host.capi.count_hits(prim.getResultBuffer(), prim.getSize())
        for name, dist, closest in [("AO", 5.0, False), ("diffuse", cam.far, True)]:
            sec = host.RayBuffer()
            tot_t, tot_r = 0.0, 0
            new = True
            rg.m_aoStartIdx = 0
            while True:
                ok, new = rg.ao(sec, prim, scene, 32, dist, new, host.FIXED_AO_SEED)
                if not ok:
                    break
                sec.setNeedClosestHit(closest)
                tracer.traceBatch(sec)
                tot_t += np.mean([tracer.traceBatch(sec) for _ in range(3)])
                tot_r += sec.getSize()
            print(f"{kernel}: {name} {tot_r / tot_t * 1e-6:.1f} Mrays/s traced ({hits * 32 / tot_t * 1e-6:.1f} counted), hits={hits}", flush=True)
    # oracle counters -> algorithmic bytes per primary ray
    res, cnt = oracle.compact_trace(nodes, woop, idx, prim.rays_host(), True, counters=True)
    m = cnt.mean(0)
    b = 32 + 16 + 64 * m[0] + 48 * m[1] + 16 * m[2] + 4 * (res[:, 0] >= 0).mean()
    print(f"primary: inner/ray {m[0]:.1f} tris/ray {m[1]:.2f} leaves/ray {m[2]:.2f} -> {b:.0f} B/ray")


if __name__ == "__main__":
    main()
