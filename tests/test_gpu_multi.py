"""Two real GPUs (skipped on a one-GPU box): NCCL broadcast of a GPU-built BVH to a replica, each rank traces a contiguous
slice of one RayBuffer (strong-scaling form), results gathered and compared with the single-GPU answer."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from ntrace_b200 import camera, capi, host, multigpu, scenes
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    host.init(rank)
    verts, tris = scenes.room(20_000, seed=7)
    scene = host.Scene(verts, tris)
    if rank == 0:
        bvh = host.HLBVHBuilder(scene)                      # built on rank 0 only
    sec = multigpu.broadcast_bvh(src=0)
    assert sec >= 0.0
    rep = host.CudaBVH(layout=host.BVHLayout_Compact); rep.resident = True
    tracer = host.CudaBVHTracer(); tracer.setBVH(rep)
    cam = camera.named_camera("conference")
    rays = host.RayBuffer()
    host.RayGen().primary(rays, cam.position, camera.nscreen_to_world(cam, 256, 192), 256, 192, cam.far)
    multigpu.trace_batch_sharded(tracer, rays, rank, world)
    lo, hi = multigpu.slice_for_rank(rays.getSize(), rank, world)
    full = multigpu.gather_results(rays.getResultBuffer()[lo:hi].contiguous(), rays.getSize(), rank, world)
    if rank == 0:
        tracer.traceBatch(rays)                              # whole batch on one GPU
        ok = bool(torch.equal(full, rays.getResultBuffer()))
        nodes_equal = True
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(ok), int(nodes_equal), hi - lo]))
    # replicas hold byte-identical BVH buffers
    n, w, i, _ = capi.bvh_download()
    t = torch.tensor([int(np.bitwise_xor.reduce(n.view(np.uint32))), int(np.bitwise_xor.reduce(w.view(np.uint32)))], device="cuda", dtype=torch.int64)
    ref = t.clone(); dist.broadcast(ref, 0)
    assert torch.equal(t, ref)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_broadcast_and_sharded_trace(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ok, _, mine = np.load(tmp_path / "ok.npy")
    assert ok == 1 and mine == 256 * 192 // 2
