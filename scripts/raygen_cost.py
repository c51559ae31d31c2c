"""dev tool: device time of nt_raygen_ao per full 1 Mi-ray diffuse batch on the bench frame, reference slot order vs the coherent order
(tile size / grid from NT_RAYGEN_TILE_R / NT_RAYGEN_CELLS), generator calls queued (nt_set_deferred(2)) so that events bracket device work."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, scenes  # noqa: E402

host.init(0)
verts, tris, cam_name = scenes.config_scene("conference")
cam = camera.named_camera(cam_name)
scene = host.Scene(verts, tris)
capi.bvh_set_collapse(1, 8)
capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, 2, 8, 0.001)
bvh = host.CudaBVH(layout=host.BVHLayout_Compact); bvh.resident = True
tracer = host.CudaBVHTracer(); tracer.setBVH(bvh)
prim = host.RayBuffer()
host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far, 0)
tracer.traceBatch(prim)
out = {}
for order in (0, 1):
    capi.raygen_set_order(order)
    rb, rg = host.RayBuffer(), host.RayGen(1 << 20)
    rg.ao(rb, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED)
    capi.synchronize()
    capi.set_deferred(2)
    capi.event_record(2)
    for _ in range(50):
        rg.m_aoStartIdx = 0
        rg.ao(rb, prim, scene, 32, cam.far, True, host.FIXED_AO_SEED)
    capi.event_record(3)
    out[order] = capi.event_elapsed(2, 3) / 50 * 1e6
    capi.set_deferred(0)
print(f"R={os.environ.get('NT_RAYGEN_TILE_R', '2')} cells={os.environ.get('NT_RAYGEN_CELLS', '16')}: reference order {out[0]:.1f} us, coherent order {out[1]:.1f} us per 1 Mi rays")
