"""Dev measurement: host-buffer nt_trace_batch wall time per 1Mi-ray batch vs chunk count (NT_E2E_CHUNKS)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes
import torch
host.init(0)
verts, tris, cam_name = scenes.config_scene("conference")
scene = host.Scene(verts, tris)
capi.bvh_set_collapse(1, 8)
bvh = host.HLBVHBuilder(scene, host.HLBVHParams(True, 2, 8, 0.001))
tracer = host.CudaBVHTracer(); tracer.setBVH(bvh)
cam = camera.named_camera(cam_name)
rg = host.RayGen(); prim = host.RayBuffer()
rg.primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
tracer.traceBatch(prim)
sec = host.RayBuffer(); rg.ao(sec, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED)
n = sec.getSize()
hr = torch.empty((n, 8), dtype=torch.float32, pin_memory=True); hr.copy_(sec.getRayBuffer())
hres = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
torch.cuda.synchronize()
dres = torch.empty((n, 4), dtype=torch.int32, device="cuda")
k = np.mean([capi.trace_batch(sec.getRayBuffer(), dres, n, True) for _ in range(5)])
for _ in range(3): capi.trace_batch(hr, hres, n, True)
t0 = time.perf_counter()
ks = [capi.trace_batch(hr, hres, n, True) for _ in range(20)]
dt = (time.perf_counter() - t0) / 20
print(f"chunks={os.environ.get('NT_E2E_CHUNKS', '8')}: device-resident kernel {k * 1e3:.3f} ms; host call {dt * 1e3:.3f} ms per {n} rays ({n / dt * 1e-6:.0f} Mrays/s), sum of chunk kernels {np.mean(ks) * 1e3:.3f} ms")
