"""BASELINE.json config 3: San Miguel stand-in (10.5 M triangles), diffuse rays sharded across N B200s with an NCCL-broadcast
replicated BVH.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multigpu_sanmiguel.py
  * rank 0 builds the BVH on its GPU, the three buffers are broadcast (NCCL over NVLink; bytes and GB/s reported)
  * STRONG scaling of one frame: every <= 1 Mi-ray diffuse batch is cut into N contiguous slices, rank g traces slice g
    (SURVEY 8e); time = max over ranks of the summed kernel seconds; also with 8 Mi-ray batches (launch cost amortised)
  * the gathered results of the first batch must equal rank 0 tracing that whole batch alone
-> gpurun_out/multigpu_sanmiguel_n<N>.json (rank 0)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ntrace_b200 import camera, capi, host, multigpu, scenes  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    host.init(local)
    verts, tris, cam_name = scenes.config_scene("sanmiguel")
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)              # every rank needs the normals for secondary-ray generation
    out = {"scene": "sanmiguel stand-in", "num_tris": int(len(tris)), "n_gpus": world}
    if rank == 0:
        capi.bvh_set_collapse(1, 8)
        lo, hi = scene.getBBox()
        capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 4, 8, 0.001)
        out["build_ms"] = min(capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, lo, hi, 4, 8, 0.001) for _ in range(3)) * 1e3
        capi.bvh_set_collapse(0, 0)
    if world > 1:
        sec = multigpu.broadcast_bvh(src=0)
        sec = min(sec, multigpu.broadcast_bvh(src=0))
        (nb, wb, ib), _ = capi.bvh_sizes()
        out["bvh_bytes"] = int(nb + wb + ib); out["broadcast_ms"] = sec * 1e3; out["broadcast_gbs"] = (nb + wb + ib) / sec * 1e-9
    bvh = host.CudaBVH(layout=4); bvh.resident = True
    tracer = host.CudaBVHTracer(); tracer.setBVH(bvh)
    prim = host.RayBuffer()
    host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, 1024, 768), 1024, 768, cam.far)
    tracer.traceBatch(prim)
    hits = capi.count_hits(prim.getResultBuffer(), prim.getSize())
    for batch_rays in (1 << 20, 1 << 23):
        gen = host.RayGen(batch_rays)
        new, total_s, first_checked, traced = True, 0.0, False, 0
        while True:
            rb = host.RayBuffer()
            ok, new = gen.ao(rb, prim, scene, 32, cam.far, new, host.FIXED_AO_SEED)
            if not ok:
                break
            rb.setNeedClosestHit(True)
            multigpu.trace_batch_sharded(tracer, rb, rank, world)                       # warm-up
            total_s += min(multigpu.trace_batch_sharded(tracer, rb, rank, world) for _ in range(3))
            traced += rb.getSize()
            if not first_checked and world > 1:
                lo_s, hi_s = multigpu.slice_for_rank(rb.getSize(), rank, world)
                full = multigpu.gather_results(rb.getResultBuffer()[lo_s:hi_s].contiguous(), rb.getSize(), rank, world)
                if rank == 0:
                    tracer.traceBatch(rb)
                    out[f"sharded_equals_single_gpu_{batch_rays}"] = bool(torch.equal(full[:, :2], rb.getResultBuffer()[:, :2]))
                first_checked = True
        t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        key = "1Mi" if batch_rays == (1 << 20) else "8Mi"
        out[f"diffuse_{key}_batches_mrays"] = hits * 32 / float(t[0]) * 1e-6
        out[f"diffuse_{key}_frame_ms"] = float(t[0]) * 1e3
        out["rays_traced"] = traced
    # the frame's batches dealt out round-robin instead of cutting every batch: rank r generates and traces batches r, r+N, ...
    # (each a full <= 1 Mi-ray launch, so no rank runs under-filled waves)
    gen = host.RayGen(1 << 20)
    new, total_s, i = True, 0.0, 0
    while True:
        rb = host.RayBuffer()
        ok, new = gen.ao(rb, prim, scene, 32, cam.far, new, host.FIXED_AO_SEED)
        if not ok:
            break
        if i % world == rank:
            rb.setNeedClosestHit(True)
            tracer.traceBatch(rb)
            total_s += min(tracer.traceBatch(rb) for _ in range(3))
        i += 1
    t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["diffuse_1Mi_batches_round_robin_mrays"] = hits * 32 / float(t[0]) * 1e-6
    out["diffuse_1Mi_batches_round_robin_frame_ms"] = float(t[0]) * 1e3
    out["batches_per_frame"] = i
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/multigpu_sanmiguel_n{world}.json", "w"), indent=1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
