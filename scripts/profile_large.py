"""One traced batch on a large scene for ncu (HBM-bound regime): python scripts/profile_large.py <sanmiguel|soup50m> <primary|diffuse>
Launches the trace kernel 3 times on the same batch (capture the last with `ncu -k regex:trace_kernel -s 2 -c 1`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402

name, rt = sys.argv[1], sys.argv[2]
host.init(0)
verts, tris, cam_name = scenes.config_scene(name)
cam = camera.named_camera(cam_name) if cam_name != "soup" else camera.look_at((1.6, 1.3, 1.1), (0.5, 0.5, 0.5), fov=60.0, near=0.001, far=10.0)
scene = host.Scene(verts, tris)
lo, hi = scene.getBBox()
capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 4, 8, 0.001)
(nb, wb, ib), _ = capi.bvh_sizes()
bvh = host.CudaBVH(layout=4); bvh.resident = True
tracer = host.CudaBVHTracer(); tracer.setBVH(bvh)
prim = host.RayBuffer()
host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
batch = prim
if rt == "diffuse":
    tracer.traceBatch(prim)
    batch = host.RayBuffer()
    gen = host.RayGen(1 << 20)
    new = True
    for _ in range(12):                                   # a batch from the middle of the image
        ok, new = gen.ao(batch, prim, scene, 32, cam.far, new, host.FIXED_AO_SEED)
    batch.setNeedClosestHit(True)
secs = [tracer.traceBatch(batch) for _ in range(3)]
print(f"{name} {rt}: {batch.getSize()} rays, BVH {(nb + wb + ib) / 1e6:.0f} MB, {batch.getSize() / min(secs) * 1e-6:.0f} Mrays/s")
