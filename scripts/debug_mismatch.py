"""Dev tool: dump rays where the GPU trace and the CPU oracle disagree (to gpurun_out/mismatch.npz)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from ntrace_b200 import camera, host, scenes

host.init(0)
verts, tris = scenes.room(20_000, seed=7, wall_frac=0.3)
cpu = oracle.CpuBVH(verts, tris, oracle.BUILDER_SPLIT, 1, 1)
nodes, woop, idx = cpu.compact()
cam = camera.named_camera("conference")
w, h = 512, 384
tracer = host.CudaBVHTracer()
tracer.setBVH(host.CudaBVH(nodes, woop, idx))
rays = host.RayBuffer()
host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, w, h), w, h, cam.far)
tracer.traceBatch(rays)
got = rays.results_host(); rh = rays.rays_host()
ref = oracle.compact_trace(nodes, woop, idx, rh, True)
mm = np.nonzero(got[:, 0] != ref[:, 0])[0]
print("mismatches", len(mm), "of", len(got))
for i in mm[:10]:
    print(i, got[i, 0], got[i, 1:].view(np.float32), ref[i, 0], ref[i, 1:2].view(np.float32))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/mismatch.npz", idx=mm, rays=rh[mm], got=got[mm], ref=ref[mm])
