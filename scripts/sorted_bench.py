"""Dev measurement: effect of RayBuffer::mortonSort (Renderer.sortRays) on AO / diffuse trace throughput, and its cost."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes

host.init(0)
import torch
verts, tris, cam_name = scenes.config_scene("conference")
scene = host.Scene(verts, tris)
bvh = host.HLBVHBuilder(scene)
tracer = host.CudaBVHTracer(); tracer.setBVH(bvh)
cam = camera.named_camera(cam_name)
rg = host.RayGen()
prim = host.RayBuffer()
rg.primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
tracer.traceBatch(prim)
for name, dist, closest in (("AO", 5.0, False), ("diffuse", cam.far, True)):
    for sort in (False, True):
        rg.m_aoStartIdx = 0; new = True
        tot_t = tot_r = 0; sort_t = 0.0
        while True:
            sec = host.RayBuffer()
            ok, new = rg.ao(sec, prim, scene, 32, dist, new, host.FIXED_AO_SEED)
            if not ok: break
            sec.setNeedClosestHit(closest)
            if sort:
                torch.cuda.synchronize(); t0 = time.perf_counter()
                sec.mortonSort()
                torch.cuda.synchronize(); sort_t += time.perf_counter() - t0
            tracer.traceBatch(sec)
            tot_t += np.mean([tracer.traceBatch(sec) for _ in range(3)]); tot_r += sec.getSize()
        print(f"{name} sort={sort}: trace {tot_r / tot_t * 1e-6:.1f} Mrays/s, sort wall {sort_t * 1e3:.1f} ms total ({tot_r / max(sort_t, 1e-9) * 1e-6:.0f} Mrays/s)", flush=True)
