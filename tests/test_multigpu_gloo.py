"""N > 1 host path on CPU: two gloo ranks exchange BVH metadata and buffers, trace disjoint contiguous slices
(the CPU oracle stands in for the per-GPU kernel here) and gather the results; the union must equal the
single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from ntrace_b200 import camera, multigpu, scenes

    verts, tris = scenes.room(4_000, seed=6)
    cam = camera.named_camera("conference")
    rays, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, 80, 60), 80, 60, cam.far)
    # rank 0 owns the BVH; replicas learn sizes from the metadata broadcast, then receive the three buffers
    if rank == 0:
        nodes, woop, idx = oracle.CpuBVH(verts, tris, oracle.BUILDER_SAH, 1, 1).compact()
        meta = torch.tensor([4, nodes.nbytes, woop.nbytes, idx.nbytes], dtype=torch.int64)
    else:
        meta = torch.zeros(4, dtype=torch.int64)
    layout, nb, wb, ib = multigpu.broadcast_meta(meta, 0)
    bufs = []
    for k, nbytes in enumerate((nb, wb, ib)):
        t = torch.from_numpy((nodes, woop, idx)[k].copy()) if rank == 0 else torch.zeros(nbytes // 4, dtype=torch.int32)
        dist.broadcast(t, 0)
        bufs.append(t.numpy())
    lo, hi = multigpu.slice_for_rank(len(rays), rank, world)
    local = oracle.compact_trace(bufs[0], bufs[1], bufs[2], rays[lo:hi], True)
    full = multigpu.gather_results(torch.from_numpy(local), len(rays), rank, world).numpy()
    if rank == 0:
        ref = oracle.compact_trace(bufs[0], bufs[1], bufs[2], rays, True)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(full, ref)), layout, len(rays), hi - lo]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_slice_and_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ok, layout, n, mine = np.load(tmp_path / "ok.npy")
    assert ok == 1 and layout == 4 and n == 4800 and mine == 2400
