"""Dev measurement: HLBVH cluster granularity (hlbvhBits) vs SAH, build time and trace throughput."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from ntrace_b200 import camera, capi, host, scenes

host.init(0)
verts, tris, cam_name = scenes.config_scene("conference")
scene = host.Scene(verts, tris)
cam = camera.named_camera(cam_name)
rg = host.RayGen(); prim = host.RayBuffer()
rg.primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
tracer = host.CudaBVHTracer()
for bits in (6, 5, 4, 3, 2, 1):
    for mode in (0, 1):
        capi.bvh_set_collapse(mode, 8)
        ts = [capi.bvh_build(1, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, bits, 8, 0.001) for _ in range(4)]
        nodes, woop, idx, _ = capi.bvh_download()
        sah = oracle.compact_sah(nodes, woop)
        bvh = host.CudaBVH(layout=4); bvh.resident = True
        tracer.setBVH(bvh)
        tracer.traceBatch(prim)
        tp = np.mean([tracer.traceBatch(prim) for _ in range(5)])
        sec = host.RayBuffer(); rg.m_aoStartIdx = 0
        rg.ao(sec, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED); sec.setNeedClosestHit(True)
        tracer.traceBatch(sec)
        td = np.mean([tracer.traceBatch(sec) for _ in range(5)])
        print(f"bits={bits} collapse={mode}: build {min(ts[1:]) * 1e3:.3f} ms, SAH {sah['sah']:.2f}, depth {sah['max_depth']}, primary {prim.getSize() / tp * 1e-6:.0f}, diffuse0 {sec.getSize() / td * 1e-6:.0f} Mrays/s", flush=True)
capi.bvh_set_collapse(0)
