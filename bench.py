#!/usr/bin/env python
"""bench.py — the reference's headline benchmark (App::runBenchmark, src/rt/App.cpp:842-1007) on B200.

One "step" = one frame's worth of the hot path over resident synthetic input on the Conference stand-in
(BASELINE.json configs[1]: ~283K triangles, 1024x768): trace the primary batch, every AO batch (any-hit,
32 spp, radius 5) and every diffuse batch (closest-hit, 32 spp), <= 1 Mi rays per batch, with a GPU-built
LBVH.  Rays counted as the reference counts them (Renderer::getTotalNumRays: w*h + hits*samples per
secondary type); Mrays/s = counted rays / device time.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference ...                      # the reference's CPU path (oracle/_ref, else the port), host cores

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1024, 768                      # App.cpp:100, config.conf:5-6
MAX_BATCH = 1 << 20                   # Renderer.cpp:45
AO_RADIUS = 5.0                       # config.conf:38
LEAF_SIZE, EPSILON = 8, 0.001         # Renderer.cpp:201-209
METRIC = "Mrays/s (primary+AO+diffuse, counted rays / trace time, Conference stand-in 283K tris, 1024x768)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs pass the count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process from a
    thread (the timed region sits in C calls that release the GIL), so even a 100 ms region gets tens of samples; an
    `nvidia-smi -lms` child started at the same time is the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, uuid=None):
        self.gpu, self.uuid, self.proc, self.path = gpu_index, uuid, None, None
        self.thread, self.stop_flag, self.nvml_sm, self.nvml_reasons, self.nvml_max = None, False, [], 0, None

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag:
            try:
                self.nvml_sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                try:
                    self.nvml_reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                except Exception:
                    self.nvml_reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
            except Exception:
                break
            time.sleep(0.004)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            handle = None
            if self.uuid:
                for u in (self.uuid, self.uuid.encode()):
                    try:
                        handle = nv.nvmlDeviceGetHandleByUUID(u)
                        break
                    except Exception:
                        handle = None
            if handle is None:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            ident = self.uuid if self.uuid else str(self.gpu)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={ident}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                for line in open(self.path):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        sm.append(float(f[1])); mx.append(float(f[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                os.unlink(self.path)
            except Exception:
                pass
        if self.nvml_sm:
            reasons |= {n for n, bit in self.REASON_BITS.items() if self.nvml_reasons & bit}
            out.update(sm_mhz=float(np.median(self.nvml_sm)), sm_max_mhz=self.nvml_max if self.nvml_max else (float(np.max(mx)) if mx else None),
                       reasons=sorted(reasons), samples=len(self.nvml_sm), source="nvml" + (f"+nvidia-smi({len(sm)})" if sm else ""))
        elif sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


# --------------------------------------------------------------------------------------------------
def make_workload():
    from ntrace_b200 import camera, scenes
    verts, tris, cam_name = scenes.config_scene("conference")
    return verts, tris, camera.named_camera(cam_name)


def gpu_uuid(torch, local):
    try:
        return "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        return None


def bytes_per_ray(cnt, hit_frac):
    """B_ray = 32 + 16 + 64 n_inner + 48 n_tri + 16 n_leaf + 4 [hit]."""
    m = cnt.mean(0)
    return float(32 + 16 + 64 * m[0] + 48 * m[1] + 16 * m[2] + 4 * hit_frac)


# --------------------------------------------------------------------------------------------------
DEFAULT_KERNEL = "b200_auto"


def lib_sha16():
    """Identity of the CUDA code being timed: sha256 over the sources libntrace_b200.so is built from (csrc/*, the C header, the nvcc flags).
    (The .so itself is not bit-reproducible across nvcc runs, so its own hash could not tie a capture to a rebuild of the same code.)"""
    from ntrace_b200 import build as nb
    return nb.source_sha16()


def load_binding(kernel):
    """What binds the kernel, measured by scripts/ncu_binding.py (ncu on the GPU box) and committed under profiles/.  The capture
    carries the sha256 of the library sources it was built from; `matches_timed_library` says whether that is the library timed here."""
    path = os.path.join(ROOT, "profiles", f"r2_binding_{kernel}.json")
    if not os.path.exists(path):
        return {"available": False, "note": f"no capture at profiles/r2_binding_{kernel}.json (run scripts/ncu_binding.py under gpurun)"}
    d = json.load(open(path))
    out = {"available": True, "source": os.path.relpath(path, ROOT), "how": d.get("how"), "captured_lib_sha16": d.get("lib_sha16"),
           "timed_lib_sha16": lib_sha16(), "l2_peak_gbs_measured": d.get("l2_peak_gbs_measured"), "hbm_peak_gbs": d.get("hbm_peak_gbs"), "per_type": {}}
    out["matches_timed_library"] = out["captured_lib_sha16"] == out["timed_lib_sha16"]
    for t, v in d.get("per_type", {}).items():
        out["per_type"][t] = {"binding": v["binding"], "frac": v["binding_frac"], "fractions_of_peak": v["fractions_of_peak"],
                              "lanes_per_instruction": v["lanes_per_instruction"], "l1_hit_rate": v["l1_hit_rate"], "l2_hit_rate": v["l2_hit_rate"],
                              "l2_gbs": v["l2_gbs"], "dram_gbs": v["dram_gbs"], "dram_bytes_per_launch": v["dram_bytes"],
                              "stall_warps_per_issue": v["stall_warps_per_issue"]}
    return out


def rank_ranges(batches, rank, world, partition="deal"):
    """This rank's share of the frame as (type, rays tensor, count, closest, offset in the type's frame buffer) launches of <= MAX_BATCH rays.
    The frame's rays of ONE type are one logical RayBuffer (the reference cuts it into <= 1 Mi-ray batches only to bound memory,
    RayGen.cpp:582-600).  partition "slices": SURVEY 8(e) literally — GPU g of G traces the contiguous slot range
    [g*ceil(n/G), min(n, (g+1)*ceil(n/G))) of that buffer.  partition "deal" (default): the buffer's batches are dealt round-robin
    (batch b goes to GPU b mod G; a type with fewer batches than GPUs is cut into contiguous slices) — same launch sizes as a single GPU,
    and every GPU gets rays from all over the image, so the slowest rank is not the one that drew the expensive half of the frame."""
    from ntrace_b200 import multigpu
    out = []
    for t in ("primary", "AO", "diffuse"):
        bs = [b for b in batches if b[0] == t]
        n = sum(b[2] for b in bs)
        if partition == "deal" and len(bs) >= world:
            base = 0
            for i, (name, rays, cnt, closest) in enumerate(bs):
                if i % world == rank:
                    out.append((name, rays, cnt, closest, base))
                base += cnt
            continue
        lo, hi = multigpu.slice_for_rank(n, rank, world)
        base = 0
        for name, rays, cnt, closest in bs:
            a, b = max(lo, base), min(hi, base + cnt)
            if b > a:
                out.append((name, rays[a - base:b - base], b - a, closest, a))
            base += cnt
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ntrace_b200 import camera, capi, host, multigpu

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host.init(local)
    dev = host.device()
    capi.set_kernel(args.kernel)
    verts, tris, cam = make_workload()
    scene = host.Scene(verts, tris)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        capi.synchronize()

    def all_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def all_sum(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    # ---- BVH: GPU build on rank 0 (timed), NCCL broadcast of the three buffers to the replicas
    build_s = []
    capi.bvh_set_collapse(args.collapse, LEAF_SIZE)
    bits = args.hlbvh_bits if args.builder == "hlbvh" else 10
    if rank == 0:
        for _ in range(1 + 5):
            build_s.append(capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, bits, LEAF_SIZE, EPSILON))
    bcast_ms = 0.0
    if world > 1:
        bcast_ms = multigpu.broadcast_bvh(src=0) * 1e3
    (node_b, woop_b, idx_b), bvh_layout = capi.bvh_sizes()
    timed_tree = capi.bvh_sah()                       # SAH of the tree that is timed (nt_bvh_sah: the reference's BVH::Stats formula, on the device)
    bvh = host.CudaBVH(layout=bvh_layout)
    bvh.resident = True
    tracer = host.CudaBVHTracer()
    tracer.setKernel(args.kernel)
    tracer.setBVH(bvh)

    # ---- resident inputs: the frame's primary rays, then every AO / diffuse batch (identical on every rank)
    capi.raygen_set_order(args.raygen_order)
    rg = host.RayGen(MAX_BATCH)
    prim = host.RayBuffer()
    rg.primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
    tracer.traceBatch(prim)
    hits = capi.count_hits(prim.getResultBuffer(), prim.getSize())
    batches = [("primary", prim.getRayBuffer(), prim.getSize(), True)]
    for name, dist_max, closest in (("AO", AO_RADIUS, False), ("diffuse", cam.far, True)):
        new = True
        while True:
            rb = host.RayBuffer()
            ok, new = rg.ao(rb, prim, scene, args.spp, dist_max, new, host.FIXED_AO_SEED)
            if not ok:
                break
            batches.append((name, rb.getRayBuffer(), rb.getSize(), closest))
    torch.cuda.synchronize()
    # what the slot order costs in the generator: one full diffuse batch generated in either order (device time, mean of 5 after a warm-up);
    # the reference times the trace kernel only (App.cpp:955-958), so this is reported beside the step, per batch of <= 1 Mi rays
    raygen_us = {}
    for order in (0, 1):
        capi.raygen_set_order(order)
        scratch, g2 = host.RayBuffer(), host.RayGen(MAX_BATCH)
        g2.ao(scratch, prim, scene, args.spp, cam.far, True, host.FIXED_AO_SEED)
        capi.synchronize()
        capi.set_deferred(2)                      # the generator is queued in this mode: the events bracket device work, not host call overhead
        capi.event_record(2)
        for _ in range(20):
            g2.m_aoStartIdx = 0
            g2.ao(scratch, prim, scene, args.spp, cam.far, True, host.FIXED_AO_SEED)
        capi.event_record(3)
        raygen_us["coherent_order" if order else "reference_order"] = capi.event_elapsed(2, 3) / 20 * 1e6
        capi.set_deferred(0)
        del scratch
    capi.raygen_set_order(args.raygen_order)
    traced = {k: sum(b[2] for b in batches if b[0] == k) for k in ("primary", "AO", "diffuse")}
    counted = {"primary": W * H, "AO": hits * args.spp, "diffuse": hits * args.spp}
    counted_step = sum(counted.values())
    ray_bytes = sum(b[2] for b in batches) * 32

    # ---- STRONG scaling (headline): the frame is fixed; every rank traces its contiguous slot range of each ray type
    mine = rank_ranges(batches, rank, world, args.partition)
    type_total = {k: traced[k] for k in traced}
    res_full = {k: torch.full((type_total[k], 4), -7, dtype=torch.int32, device=dev) for k in type_total}     # results in place, per type
    res_alt = [torch.empty((MAX_BATCH, 4), dtype=torch.int32, device=dev) for _ in range(2)]

    CLOSEST = {"primary": True, "AO": False, "diffuse": True}

    def frame_launches(share, dst_of):
        """--submission frame: this rank's share of a ray type in ONE persistent launch (nt_trace_batches); a single batch of camera rays
        keeps the plain call (b200_auto recognises the buffer its own generator wrote)."""
        for name in ("primary", "AO", "diffuse"):
            grp = [(rays, n, off) for (nm, rays, n, closest, off) in share if nm == name]
            if not grp:
                continue
            if name == "primary" and len(grp) == 1:
                capi.trace_batch(grp[0][0], dst_of(name, grp[0][2], grp[0][1]), grp[0][1], True)
            else:
                capi.trace_batches([g_[0] for g_ in grp], [dst_of(name, g_[2], g_[1]) for g_ in grp], [g_[1] for g_ in grp], CLOSEST[name])

    def step_strong(keep=False, submission=None):
        if (submission or args.submission) == "frame":
            frame_launches(mine, lambda name, off, n: res_full[name][off:off + n])        # every batch needs its own result range
            return
        for i, (name, rays, n, closest, off) in enumerate(mine):
            dst = res_full[name][off:off + n] if keep else res_alt[i & 1][:n]
            capi.trace_batch(rays, dst, n, closest)

    capi.set_deferred(2 if args.overlap else 1)
    for _ in range(args.warmup):
        step_strong()
    barrier()
    sampler = ClockSampler(local, gpu_uuid(torch, local))
    sampler.start()
    l0 = capi.launch_count()
    capi.event_record(0)
    for _ in range(args.steps):
        step_strong()
    capi.event_record(1)
    sec = capi.event_elapsed(0, 1)
    launches = capi.launch_count() - l0
    barrier()
    clocks = sampler.stop()

    if args.profile:
        capi.set_deferred(False)
        if rank == 0:
            print(f"profile run: {counted_step * args.steps / sec * 1e-6:.1f} Mrays/s (number taken under a profiler is not a bench value)")
        return None

    # one launch per <= 1 Mi-ray batch (the reference's loop), queued on two kernel streams (the headline before nt_trace_batches) ...
    capi.set_deferred(2)
    step_strong(submission="batches")
    barrier()
    capi.event_record(0)
    for _ in range(args.steps):
        step_strong(submission="batches")
    capi.event_record(1)
    sec_two_streams = capi.event_elapsed(0, 1)
    barrier()
    # ... and on one stream (no tail overlap)
    capi.set_deferred(1)
    step_strong(submission="batches")
    barrier()
    capi.event_record(0)
    for _ in range(args.steps):
        step_strong(submission="batches")
    capi.event_record(1)
    sec_serial = capi.event_elapsed(0, 1)
    barrier()

    # ---- sharded results vs the single-GPU answer, inside the run (results kept in place this time)
    capi.set_deferred(2 if args.overlap else 1)
    step_strong(keep=True)
    capi.synchronize()
    capi.set_deferred(False)
    sharded_ok, gathered_ok = True, None
    single = {}
    for name, rays, n, closest in batches:                       # every rank holds the whole frame: its own single-GPU answer
        r = torch.empty((n, 4), dtype=torch.int32, device=dev)
        capi.trace_batch(rays, r, n, closest)
        single.setdefault(name, []).append(r)
    single = {k: torch.cat(v) for k, v in single.items()}
    for name, rays, n, closest, off in mine:
        if not torch.equal(res_full[name][off:off + n], single[name][off:off + n]):
            sharded_ok = False
    if world > 1:
        ok = torch.tensor([1.0 if sharded_ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_ok = bool(ok.item() > 0.5)
        # and literally gathered: every rank's diffuse results, in place in a frame-sized buffer that is zero elsewhere, summed over NCCL
        # (each slot is written by exactly one rank), against this rank's own single-GPU trace of the whole frame
        full = torch.zeros_like(res_full["diffuse"])
        for name, rays, n, closest, off in mine:
            if name == "diffuse":
                full[off:off + n] = res_full["diffuse"][off:off + n]
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        gathered_ok = bool(torch.equal(full, single["diffuse"]))
        g = torch.tensor([1.0 if gathered_ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
        gathered_ok = bool(g.item() > 0.5)

    # ---- WEAK scaling beside it: every rank traces a whole frame (per-GPU work fixed)
    whole = []
    offs = {"primary": 0, "AO": 0, "diffuse": 0}
    for name, rays, n, closest in batches:
        whole.append((name, rays, n, closest, offs[name]))
        offs[name] += n

    def step_weak():
        if args.submission == "frame":
            frame_launches(whole, lambda name, off, n: res_full[name][off:off + n])
            return
        for i, (_, rays, n, closest) in enumerate(batches):
            capi.trace_batch(rays, res_alt[i & 1][:n], n, closest)

    weak_sec = None
    if world > 1:
        capi.set_deferred(2 if args.overlap else 1)
        step_weak()
        barrier()
        capi.event_record(0)
        for _ in range(max(1, args.steps // 2)):
            step_weak()
        capi.event_record(1)
        weak_sec = capi.event_elapsed(0, 1) / max(1, args.steps // 2)
        barrier()
        capi.set_deferred(False)

    # ---- per-type kernel time (synchronous calls, CUDA events around each launch: the reference's accounting), rank 0's whole frame
    # three passes over the frame, mean per launch (the primary type is ONE launch per frame: a single sample of 0.2 ms is noise)
    type_sec = {"primary": 0.0, "AO": 0.0, "diffuse": 0.0}
    res_dev = res_alt[0]
    TYPE_PASSES = 3
    for _ in range(TYPE_PASSES):
        for name, rays, n, closest in batches:
            type_sec[name] += capi.trace_batch(rays, res_dev[:n], n, closest) / TYPE_PASSES

    # ---- the same per type with the whole frame's batches in ONE launch (nt_trace_batches, synchronous: events around the launch)
    frame_launch_mrays = {}
    for name in ("AO", "diffuse"):
        grp = [(rays, n, off) for (nm, rays, n, closest, off) in whole if nm == name]
        if not grp:
            continue
        lists = ([g_[0] for g_ in grp], [res_full[name][g_[2]:g_[2] + g_[1]] for g_ in grp], [g_[1] for g_ in grp])
        capi.trace_batches(*lists, CLOSEST[name])
        t = float(np.mean([capi.trace_batches(*lists, CLOSEST[name]) for _ in range(3)]))
        frame_launch_mrays[name] = counted[name] / t * 1e-6

    # ---- e2e: this rank's share of the frame through the C ABI with HOST buffers (pinned), H2D of the rays and D2H of the results
    # inside the timed region, every step
    prev_affinity, numa_node = multigpu.bind_host_to_gpu(local)      # pinned buffers land on the GPU's NUMA node
    host_batches = []
    for name, rays, n, closest, off in mine:
        hb = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        hb.copy_(rays)
        host_batches.append((hb, n, closest))
    my_rays = sum(b[1] for b in host_batches)
    torch.cuda.synchronize()
    e2e_steps = max(1, args.steps)
    NSLOT = 3
    res_slots = [torch.empty((MAX_BATCH, 4), dtype=torch.int32, pin_memory=True) for _ in range(NSLOT)]

    def sync_frame():
        for hb, n, closest in host_batches:
            capi.trace_batch(hb, res_slots[0], n, closest)

    def pipelined_frame():
        # nt_trace_batch_async keeps NSLOT independent batches of the frame in flight (H2D of batch i+1, traversal of batch i and
        # D2H of batch i-1 overlap); every batch's rays still cross PCIe in and its results out
        for i, (hb, n, closest) in enumerate(host_batches):
            s = i % NSLOT
            capi.trace_wait(s)
            capi.trace_batch_async(hb, res_slots[s], n, closest, s)
        for s in range(NSLOT):
            capi.trace_wait(s)

    sync_frame()                                  # warm-up passes (staging buffers grow here)
    pipelined_frame()
    barrier()
    capi.event_record(2)
    for _ in range(min(e2e_steps, 3)):
        sync_frame()
    capi.event_record(3)
    e2e_sync_sec = capi.event_elapsed(2, 3) / min(e2e_steps, 3)
    barrier()
    capi.event_record(2)
    for _ in range(e2e_steps):
        pipelined_frame()
    capi.event_record(3)
    e2e_sec = capi.event_elapsed(2, 3) / e2e_steps
    barrier()
    # what the host link can deliver with every rank copying at once: pinned H2D and D2H of 256 MiB per rank, both directions busy
    pcie = pcie_probe(torch, barrier)
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)                        # the CPU baseline leg uses every core again

    # ---- max over ranks
    sec, sec_serial, sec_two_streams, e2e_sec, e2e_sync_sec = all_max([sec, sec_serial, sec_two_streams, e2e_sec, e2e_sync_sec])
    launches_all = int(all_sum([launches])[0])
    h2d_all, d2h_all, bi_h2d_all, bi_d2h_all = all_sum([pcie["h2d_gbs"], pcie["d2h_gbs"], pcie["bidir_h2d_gbs"], pcie["bidir_d2h_gbs"]])
    # rays/s the host link allows for this traffic mix (32 B in, 16 B out per ray): each direction alone, and their sum when both are busy
    pcie_roof_mrays = min(h2d_all / 32, d2h_all / 16, (bi_h2d_all + bi_d2h_all) / 48) * 1e3
    weak_value = None
    if weak_sec is not None:
        weak_value = counted_step * world / all_max([weak_sec])[0] * 1e-6

    value = counted_step * args.steps / sec * 1e-6                      # the frame is fixed: whole-job rays / slowest rank's time
    e2e_value = counted_step / e2e_sec * 1e-6

    # ---- rank 0: the checker legs (oracle) on the timed frame, while the conference BVH is still the resident one
    cpu = ref_gpu_rows = None
    if rank == 0:
        gpu_bvh = capi.bvh_download()[:3]
        # strided ~500K-ray sample of every ray type, taken across all batches of the timed frame, with the GPU results of those rays
        samples, gpu_res = [], {}
        for k, closest in (("primary", True), ("AO", False), ("diffuse", True)):
            stride = max(1, traced[k] // 500_000)
            rs, gs, off = [], [], 0
            for b in batches:
                if b[0] == k:
                    rs.append(b[1][::stride])
                    gs.append(single[k][off:off + b[2]][::stride])
                    off += b[2]
            samples.append((k, torch.cat(rs).cpu().numpy(), closest))
            gpu_res[k] = torch.cat(gs).cpu().numpy()
        cpu = cpu_baseline_leg(verts, tris, cam, args, gpu_bvh=gpu_bvh, sample_batches=samples, gpu_results=gpu_res, kernel=args.kernel)
        ref_gpu_rows = reference_gpu_leg(torch, capi, batches) if args.reference_gpu else None
        capi.set_kernel(args.kernel)

    config3 = None
    if world > 1 and args.config3:
        config3 = config3_leg(torch, dist, capi, host, multigpu, camera, rank, world, dev, args, barrier, all_max)
        capi.set_kernel(args.kernel)

    # ---- builder figures next to the trace figures (rank 0, single-GPU runs): the bench tree, the plain LBVH of the same scene and a
    # 10 M-triangle soup (the north_star's build target), each with the algorithmic-bytes fraction of the measured HBM peak
    build_block = None
    if rank == 0 and world == 1 and args.build_leg:
        build_block = build_leg(torch, capi, scene, bits, args)

    out = None
    if rank == 0:
        peak, peak_src = load_peaks()
        # algorithmic bytes of one step = sum over ray types of mean B_ray (oracle counters on the
        # GPU-built BVH, ~500K-ray strided sample per type) x rays traced of that type
        alg_bytes_step = sum(cpu["bytes_per_ray"][k] * traced[k] for k in traced)
        n_launch_step = len(batches)
        serial_launch_s = sum(type_sec.values()) / n_launch_step          # one launch at a time, CUDA events around each launch
        achieved = alg_bytes_step / n_launch_step / serial_launch_s * 1e-9
        binding = load_binding(args.kernel)
        traffic = None
        if binding.get("available"):
            pt = binding["per_type"]
            traffic = sum(pt[k]["dram_bytes_per_launch"] * sum(1 for b in batches if b[0] == k) for k in pt) / n_launch_step
            # the fraction that grades the kernel: the most-utilised hardware unit, weighted by where the step spends its time
            w = {k: type_sec[k] / sum(type_sec.values()) for k in type_sec}
            binding["step_weighted"] = {u: sum(w[k] * pt[k]["fractions_of_peak"][u] for k in pt) for u in next(iter(pt.values()))["fractions_of_peak"]}
            top = max(binding["step_weighted"], key=binding["step_weighted"].get)
            binding["binding"], binding["frac"] = top, binding["step_weighted"][top]
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "conference stand-in room(283000, seed=2): primary + AO(32spp, r=5, any-hit) + diffuse(32spp, closest-hit), "
                                   "1024x768, <= 1 Mi rays per batch (RayBuffer), GPU %s leaf 8%s" % ("HLBVH(hlbvhBits %d)" % args.hlbvh_bits if args.builder == "hlbvh" else "LBVH", ", SAH-guided collapse" if args.collapse else ""),
                       "rays_traced_per_step": int(sum(traced.values())), "rays_counted_per_step": int(counted_step),
                       "launches_per_step_per_gpu": launches // max(1, args.steps), "batches_per_step_per_gpu": len(mine), "kernel": args.kernel,
                       "ray_order": ("nt_raygen_set_order(1): the generator writes each tile of <= 1024 secondary rays (32 neighbouring hit points x 32 samples) in "
                                     "direction-cell order (same rays and ids, slot permutation in idToSlot / slotToID, no extra pass; detail.raygen_us_per_batch)"
                                     if args.raygen_order else "the reference generator's slot order (slot = id)"),
                       "submission": (("nt_trace_batches: a rank's share of each ray type (its <= 1 Mi-ray batches, each with its own ray and result buffer) is ONE persistent "
                                       "launch, so only one ramp-up / drain per type remains; detail.per_batch_two_streams_value / one_stream_value are the same step with one "
                                       "launch per batch" if args.submission == "frame" else
                                       "one launch per <= 1 Mi-ray batch (the reference's loop), ") +
                                      ("; nt_set_deferred(2): launches are queued, consecutive per-batch launches alternate between two kernel streams"
                                       if args.overlap else "; nt_set_deferred(1): launches are queued on one stream")),
                       "l2": "inputs exceed L2: %.0f MB of rays per step stream from HBM; the %.0f MB BVH is reused within a frame by design"
                             % (ray_bytes / 1e6, (node_b + woop_b + idx_b) / 1e6),
                       "parallelism": ("the frame is fixed (strong scaling): BVH built on rank 0 and replicated by NCCL broadcast; each ray type's frame buffer is split over the GPUs "
                                       + ("by dealing its <= 1 Mi-ray batches round-robin (a type with fewer batches than GPUs is cut into contiguous slices)" if args.partition == "deal"
                                          else "into contiguous slot ranges [g*ceil(n/G), (g+1)*ceil(n/G)) (SURVEY 8e)")
                                       + "; no collective on the ray path") if world > 1 else "single GPU"},
            "detail": {"primary_mrays": counted["primary"] / type_sec["primary"] * 1e-6, "ao_mrays": counted["AO"] / type_sec["AO"] * 1e-6,
                       "diffuse_mrays": counted["diffuse"] / type_sec["diffuse"] * 1e-6,
                       # the three figures above: one synchronous launch per <= 1 Mi-ray batch (the reference's accounting), ramp-up and drain of
                       # every launch included; the same rays with the frame's batches of a type in one launch:
                       "ao_mrays_frame_launch": frame_launch_mrays.get("AO"), "diffuse_mrays_frame_launch": frame_launch_mrays.get("diffuse"),
                       "build_ms": float(np.mean(build_s[1:]) * 1e3), "build_mtris": len(tris) / float(np.mean(build_s[1:])) * 1e-6,
                       "timed_tree": timed_tree,
                       "bvh_broadcast_ms": bcast_ms, "bvh_broadcast_via": "nt_bvh_broadcast (C ABI, NCCL bound by dlopen)" if multigpu._comm_ready else ("torch.distributed" if world > 1 else None),
                       "primary_hits": int(hits), "raygen_us_per_batch": raygen_us,
                       # the same step if the extra device time the coherent order costs in the generator were charged to it (the reference's
                       # accounting, App.cpp:955-958, times the trace kernels only)
                       "value_charging_raygen_delta": (counted_step / (sec / args.steps + (len(mine) - 1) * max(0.0, raygen_us["coherent_order"] - raygen_us["reference_order"]) * 1e-6) * 1e-6
                                                       if args.raygen_order else value),
                       "per_batch_two_streams_value": counted_step * args.steps / sec_two_streams * 1e-6,
                       "one_stream_value": counted_step * args.steps / sec_serial * 1e-6,
                       "overlap_gain": sec_serial / sec},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(ray_bytes), "d2h_bytes_per_step": int(ray_bytes // 2),
                    "steps": e2e_steps, "host_numa_node": numa_node,
                    "api": "nt_trace_batch_async/nt_trace_wait, 3 batches in flight per rank, pinned host rays in and results out every batch",
                    "sync_value": counted_step / e2e_sync_sec * 1e-6,
                    "sync_api": "nt_trace_batch (one synchronous call per batch, zero-copy pinned buffers)",
                    "pcie_roof_gbs": {"h2d_alone_all_ranks": h2d_all, "d2h_alone_all_ranks": d2h_all, "bidir_h2d_all_ranks": bi_h2d_all, "bidir_d2h_all_ranks": bi_d2h_all,
                                      "how": "every rank copies 256 MiB pinned<->device at the same time: host->device alone, device->host alone, both at once (two streams); best of 2, summed over ranks"},
                    "pcie_roof_mrays": pcie_roof_mrays,
                    "roof_rule": "min(h2d_alone / 32 B, d2h_alone / 16 B, (bidir_h2d + bidir_d2h) / 48 B) per ray",
                    "h2d_gbs_achieved": ray_bytes / e2e_sec * 1e-9, "d2h_gbs_achieved": ray_bytes / 2 / e2e_sec * 1e-9,
                    "frac_of_pcie_roof": e2e_value / pcie_roof_mrays if pcie_roof_mrays > 0 else None},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "parity": cpu["parity"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "note": "SURVEY 8(d) figure: algorithmic bytes (nodes + triangles fetched per ray, oracle-counted on the timed BVH) over the SERIALISED launch time "
                                 "(one launch at a time, CUDA events around each launch; the overlapped submission's gain is detail.overlap_gain). The BVH is L2/L1-resident, "
                                 "so these bytes never reach HBM and frac can exceed 1: the unit that actually binds, measured by ncu on this library, is in `binding`",
                         "bytes_per_ray": cpu["bytes_per_ray"],
                         "binding": binding},
            "strong": {"sharded_equals_single_gpu": sharded_ok, "gathered_equals_single_gpu": gathered_ok,
                       "how": "every rank re-traces the whole frame alone and compares its slices bit for bit (all ranks must agree); the diffuse results are also "
                              "all-gathered over NCCL and compared with the single-GPU trace" if world > 1 else "single GPU: the deferred, overlapped submission against synchronous calls"},
            "weak": None if weak_value is None else {"value": weak_value, "unit": "Mrays/s", "what": "every rank traces a whole frame (per-GPU work fixed)"},
            "config3": config3,
            "build": build_block,
            "cpu_baseline": cpu["baseline"],
            "reference_gpu": ref_gpu_rows,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        if multigpu._comm_ready:
            capi.comm_destroy()
        dist.destroy_process_group()
    if out is not None and not (out["parity"]["ok"] and sharded_ok and gathered_ok is not False):
        print("bench.py: PARITY FAILURE — the line above must not be used", file=sys.stderr)
        raise SystemExit(3)
    return out


def pcie_probe(torch, barrier, nbytes=256 << 20):
    """Pinned host <-> device copy bandwidth of THIS rank while every other rank does the same (GB/s): host->device alone, device->host
    alone, and both directions at once (two streams)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {"h2d_gbs": 0.0, "d2h_gbs": 0.0, "bidir_h2d_gbs": 0.0, "bidir_d2h_gbs": 0.0}
    for mode in ("h2d", "d2h", "bidir"):
        for it in range(3):
            barrier()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if mode != "d2h":
                with torch.cuda.stream(s1):
                    e[0].record()
                    for _ in range(4):
                        d_in.copy_(h_in, non_blocking=True)
                    e[1].record()
            if mode != "h2d":
                with torch.cuda.stream(s2):
                    e[2].record()
                    for _ in range(4):
                        h_out.copy_(d_out, non_blocking=True)
                    e[3].record()
            torch.cuda.synchronize()
            if not it:
                continue
            if mode != "d2h":
                k = "h2d_gbs" if mode == "h2d" else "bidir_h2d_gbs"
                out[k] = max(out[k], 4 * nbytes / (e[0].elapsed_time(e[1]) * 1e-3) * 1e-9)
            if mode != "h2d":
                k = "d2h_gbs" if mode == "d2h" else "bidir_d2h_gbs"
                out[k] = max(out[k], 4 * nbytes / (e[2].elapsed_time(e[3]) * 1e-3) * 1e-9)
    return out


def config3_leg(torch, dist, capi, host, multigpu, camera, rank, world, dev, args, barrier, all_max):
    """BASELINE.json configs[3]: San Miguel stand-in (10.5 M triangles), one frame of diffuse rays (32 spp) strong-scaled over the ranks:
    rank 0 builds, the three buffers are broadcast over NCCL (timed), rank g traces its contiguous slot range of the frame's diffuse
    RayBuffer in <= 1 Mi-ray launches, the results are gathered and compared with rank 0 tracing the whole frame alone."""
    from ntrace_b200 import scenes
    verts, tris, cam_name = scenes.config_scene("sanmiguel")
    cam = camera.named_camera(cam_name)
    scene = host.Scene(verts, tris)
    out = {"workload": "San Miguel stand-in room(10500000, seed=4), 1024x768, diffuse 32 spp closest-hit, GPU HLBVH(4) leaf 8 + SAH collapse", "num_tris": int(len(tris))}
    capi.bvh_set_collapse(1, LEAF_SIZE)
    if rank == 0:
        capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, 4, LEAF_SIZE, EPSILON)
        out["build_ms"] = min(capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax, 4, LEAF_SIZE, EPSILON) for _ in range(2)) * 1e3
    sec_b = multigpu.broadcast_bvh(src=0)
    sec_b = min(sec_b, multigpu.broadcast_bvh(src=0))
    (nb, wb, ib), layout = capi.bvh_sizes()
    out.update(bvh_bytes=int(nb + wb + ib), broadcast_ms=sec_b * 1e3, broadcast_gbs=(nb + wb + ib) / sec_b * 1e-9)
    bvh = host.CudaBVH(layout=layout)
    bvh.resident = True
    tracer = host.CudaBVHTracer()
    tracer.setKernel(args.kernel)
    tracer.setBVH(bvh)
    prim = host.RayBuffer()
    host.RayGen().primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far)
    tracer.traceBatch(prim)
    hits = capi.count_hits(prim.getResultBuffer(), prim.getSize())
    rg = host.RayGen(MAX_BATCH)
    batches, new = [], True
    while True:
        rb = host.RayBuffer()
        ok, new = rg.ao(rb, prim, scene, args.spp, cam.far, new, host.FIXED_AO_SEED)
        if not ok:
            break
        batches.append(("diffuse", rb.getRayBuffer(), rb.getSize(), True))
    n_total = sum(b[2] for b in batches)
    mine = rank_ranges(batches, rank, world, args.partition)
    res_full = torch.full((n_total, 4), -7, dtype=torch.int32, device=dev)

    def frame(parts, keep):
        if args.submission == "frame" and len(parts) > 1:            # this rank's share of the frame in one persistent launch
            capi.trace_batches([p_[1] for p_ in parts], [res_full[p_[4]:p_[4] + p_[2]] for p_ in parts], [p_[2] for p_ in parts], True)
            return
        for i, (_, rays, n, closest, off) in enumerate(parts):
            capi.trace_batch(rays, res_full[off:off + n], n, closest)

    def timed(parts, reps=3):
        capi.set_deferred(2 if args.overlap else 1)
        frame(parts, True)
        barrier()
        best = 1e30
        for _ in range(reps):
            capi.event_record(4)
            frame(parts, True)
            capi.event_record(5)
            best = min(best, capi.event_elapsed(4, 5))
            barrier()
        capi.set_deferred(False)
        return best

    t_sharded = all_max([timed(mine)])[0]
    out["mrays_sharded"] = hits * args.spp / t_sharded * 1e-6
    out["frame_ms_sharded"] = t_sharded * 1e3
    out["batches_per_gpu"] = len(mine)
    out["submission"] = args.submission
    out["partition"] = args.partition
    gathered = torch.zeros_like(res_full)                      # every slot is written by exactly one rank: the NCCL sum assembles the frame
    for _, rays, n, closest, off in mine:
        gathered[off:off + n] = res_full[off:off + n]
    dist.all_reduce(gathered, op=dist.ReduceOp.SUM)
    # the single-GPU answer and time, same run: rank 0 traces the whole frame alone (the other ranks run the same barriers with nothing to trace)
    t_single = timed(rank_ranges(batches, 0, 1, args.partition) if rank == 0 else [])
    same = True
    if rank == 0:
        same = bool(torch.equal(gathered, res_full))
        out["mrays_single_gpu_same_run"] = hits * args.spp / t_single * 1e-6
        out["frame_ms_single_gpu"] = t_single * 1e3
    ok = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out["sharded_equals_single_gpu"] = bool(ok.item() > 0.5)
    out["rays_traced"] = int(n_total)
    return out


# --------------------------------------------------------------------------------------------------
LBVH_BYTES_PER_TRI = 350.0            # 330 + 150 / leafSize algorithmic bytes per triangle at leaf size 8


def build_leg(torch, capi, scene, bits, args):
    """GPU build times (CUDA events around the whole pipeline, mean of 5 after one warm-up) and what they are of the HBM roofline."""
    from ntrace_b200 import scenes
    peak, _ = load_peaks()

    def timed(builder, verts, tris, lo, hi, b, collapse):
        capi.bvh_set_collapse(collapse, LEAF_SIZE)
        ts = [capi.bvh_build(builder, verts, tris, lo, hi, b, LEAF_SIZE, EPSILON) for _ in range(6)][1:]
        n = int(tris.shape[0])
        row = {"ms": float(np.mean(ts) * 1e3), "ms_best": float(np.min(ts) * 1e3), "mtris": n / float(np.mean(ts)) * 1e-6}
        if builder == capi.BUILDER_LBVH:
            row["algorithmic_gbs"] = n * LBVH_BYTES_PER_TRI / float(np.mean(ts)) * 1e-9
            row["frac_of_hbm_peak"] = row["algorithmic_gbs"] / peak
        return row

    out = {"bytes_per_tri_lbvh": LBVH_BYTES_PER_TRI, "hbm_peak_gbs": peak,
           "how": "nt_bvh_build, CUDA events around the whole pipeline (device-resident vertices / indices), mean of 5 builds after one warm-up; "
                  "frac_of_hbm_peak = triangles x algorithmic bytes per triangle (DESIGN.md 3.2) / time / measured HBM peak"}
    v, t, lo, hi = scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax
    out["conference_283k_bench_tree"] = timed(capi.BUILDER_HLBVH, v, t, lo, hi, bits, args.collapse)
    out["conference_283k_hlbvh4_reference_renderer"] = timed(capi.BUILDER_HLBVH, v, t, lo, hi, 4, 0)
    out["conference_283k_lbvh"] = timed(capi.BUILDER_LBVH, v, t, lo, hi, 10, 0)
    sv, st = scenes.soup_uniform(10_000_000, 5)
    slo, shi = scenes.bbox(sv)
    dv, dt = torch.from_numpy(sv).cuda(), torch.from_numpy(st).cuda()
    out["soup_10m_lbvh"] = timed(capi.BUILDER_LBVH, dv, dt, slo, shi, 10, 0)
    out["soup_10m_hlbvh4"] = timed(capi.BUILDER_HLBVH, dv, dt, slo, shi, 4, 0)
    del dv, dt
    capi.bvh_set_collapse(args.collapse, LEAF_SIZE)
    return out


def cpu_baseline_leg(verts, tris, cam, args, gpu_bvh=None, sample_batches=None, gpu_results=None, kernel=None):
    """The ONLY place bench.py touches the oracle: (a) times the reference's CPU path (restated SplitBVHBuilder +
    BVH::trace) on a bounded sample, (b) counts nodes/triangles per ray on the GPU-built BVH for the roofline."""
    import oracle
    from ntrace_b200 import camera
    threads = host_threads()
    t0 = time.time()
    cpu = oracle.CpuBVH(verts, tris, oracle.BUILDER_SPLIT, 1, 1, 1.0e-5)
    build_s = time.time() - t0
    st = cpu.stats()
    if sample_batches is None:
        rays, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
        res = cpu.trace(rays, True, nthreads=threads)
        normals = oracle.tri_normals(verts, tris)
        n_in = MAX_BATCH // args.spp
        ao, _, _ = oracle.raygen_ao(rays, res, normals, 0, n_in, args.spp, AO_RADIUS, 0x9E3779B9)
        df, _, _ = oracle.raygen_ao(rays, res, normals, 0, n_in, args.spp, cam.far, 0x9E3779B9)
        sample_batches = [("primary", rays, True), ("AO", ao, False), ("diffuse", df, True)]
    total_rays, total_s = 0, 0.0
    for _, r, closest in sample_batches:
        t0 = time.time()
        cpu.trace(r, closest, nthreads=threads)
        total_s += time.time() - t0
        total_rays += len(r)
    bpr = {}
    parity = {"ok": True, "checker": "oracle.compact_trace (restated CudaBVH::trace, pinned to the reference) on the BVH downloaded from the GPU, strided sample of the TIMED frame",
              "tolerance": "north_star: ids identical on >= 99.99 % of rays, mismatches only where t agrees within 1e-4 rel; t within 1e-5 rel; any-hit rays: hit/miss flag"}
    if gpu_bvh is not None:
        for name, r, closest in sample_batches:
            res, cnt = oracle.compact_trace(gpu_bvh[0], gpu_bvh[1], gpu_bvh[2], r, closest, counters=True, nthreads=threads)
            bpr[name] = bytes_per_ray(cnt, float((res[:, 0] >= 0).mean()))
            if gpu_results is not None:
                g = gpu_results[name]
                hit_r, hit_g = res[:, 0] >= 0, g[:, 0] >= 0
                row = {"rays": int(len(r)), "flag_match": float((hit_r == hit_g).mean()), "id_match": float((res[:, 0] == g[:, 0]).mean())}
                ok = row["flag_match"] >= 0.9999
                if closest:
                    tr, tg = res[:, 1].view(np.float32), g[:, 1].view(np.float32)
                    both = hit_r & hit_g
                    rel = np.abs(tr - tg)[both] / np.maximum(np.abs(tr[both]), 1e-30)
                    same = res[:, 0] == g[:, 0]
                    row["max_rel_t"] = float(rel.max()) if both.any() else 0.0
                    mm = (~same) & both
                    row["id_mismatch_max_rel_t"] = float((np.abs(tr - tg)[mm] / np.maximum(np.abs(tr[mm]), 1e-30)).max()) if mm.any() else 0.0
                    ok = ok and row["id_match"] >= 0.9999 and row["max_rel_t"] <= 1e-4 and float(rel[same[both]].max() if same[both].any() else 0.0) <= 1e-5
                row["ok"] = bool(ok)
                parity[name] = row
                parity["ok"] = parity["ok"] and bool(ok)
    baseline = {"value": total_rays / total_s * 1e-6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                "sample": "restated BVH::trace (BVH.cpp:90-186) on a restated SplitBVH (alpha 1e-5, leaf 1/1) of the same scene, OpenMP over rays, "
                          + ", ".join(f"{len(r)} {n}" for n, r, _ in sample_batches) + " rays",
                "build_s_1thread": build_s, "build_mtris_1thread": len(tris) / build_s * 1e-6, "splitbvh_sah": st.sah}
    return {"baseline": baseline, "bytes_per_ray": bpr, "parity": parity}


def reference_gpu_leg(torch, capi, batches):
    """Baseline beside the CPU one: the reference's OWN traversal kernels recompiled for sm_100a (oracle/ref_gpu.py,
    test infrastructure) on the bench's BVH and on the first batch of each ray type, kernel time only.  `as_shipped` is the
    launch the reference's host would do (fermi: one thread per ray; kepler: its hard-coded 720 persistent warps),
    `resized_launch` gives the persistent kernel 148 SMs x {16..64} warps with its code untouched.  None when the
    recompiled kernels are not present."""
    try:
        from oracle import ref_gpu
        if not all(ref_gpu.available(k) for k in ref_gpu.KERNELS):
            return None
        first = {}
        for name, rays, n, closest in batches:
            first.setdefault(name, (rays, n, closest))
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        rows = {"note": "reference kernels (src/rt/kernels/*.cu) compiled unmodified for sm_100a with the reference's -use_fast_math; same BVH, first batch of each ray type; Mrays/s = rays traced / kernel time"}
        for kernel, layout in (("fermi_speculative_while_while", 4), ("kepler_dynamic_fetch", 5)):
            capi.bvh_convert(layout)
            nodes, woop, idx, _ = capi.bvh_download()
            d = [torch.from_numpy(a).cuda() for a in (nodes, woop, idx)]
            row = {}
            for name, (rays, n, closest) in first.items():
                res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
                ms, cfg = ref_gpu.trace(kernel, rays[:n], res, d[0], d[1], d[2], any_hit=not closest, repeats=5)
                row[name + "_as_shipped"] = n / ms * 1e-3
                if cfg["usePersistentThreads"]:
                    row[name + "_resized_launch"] = max(n / ref_gpu.trace(kernel, rays[:n], res, d[0], d[1], d[2], any_hit=not closest,
                                                                          desired_warps=sms * w, repeats=3)[0] * 1e-3 for w in (16, 32, 48, 64))
            rows[kernel] = row
        capi.bvh_convert(4)
        return rows
    except Exception as e:                       # a baseline leg must never take the bench line down
        try:
            capi.bvh_convert(4)
        except Exception:
            pass
        return {"error": str(e)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, same metric/config; rank 0
    only.  When oracle/_ref/libref.so exists (the reference's SplitBVHBuilder + CudaBVH::trace — the tracer its CPURenderer
    calls, CPURenderer.cpp:108-124 — compiled unmodified, oracle/Makefile) that is what runs (kind "reference"); otherwise
    the restated port (kind "port").  Rays come from the restated ray generator in both cases (input data, untimed)."""
    if env_int("RANK", 0) != 0:
        return
    import oracle
    from oracle import ref
    from ntrace_b200 import camera
    verts, tris, cam = make_workload()
    threads = host_threads()
    use_ref = ref.available() and os.environ.get("NT_BENCH_REFERENCE_PORT", "0") != "1"
    if use_ref:
        cpu = ref.RefBVH(verts, tris, split=True, min_leaf=1, max_leaf=1, split_alpha=1.0e-5)     # Renderer.cpp:88-89 leaf prefs

        def trace(r, closest):
            return cpu.compact_trace(r, closest, nthreads=threads)
        what = "the reference's own SplitBVHBuilder + CudaBVH::trace (compiled unmodified: oracle/_ref), one CudaBVH per thread, rays split evenly"
    else:
        cpu = oracle.CpuBVH(verts, tris, oracle.BUILDER_SPLIT, 1, 1, 1.0e-5)

        def trace(r, closest):
            return cpu.trace(r, closest, nthreads=threads)
        what = "restated SplitBVH + BVH::trace, OpenMP"
    rays, _, _ = oracle.raygen_primary(cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0)
    res = trace(rays, True)
    hits = oracle.count_hits(res)
    normals = oracle.tri_normals(verts, tris)
    n_in = MAX_BATCH // args.spp                     # one secondary batch of each type per step (bounded sample)
    ao, _, _ = oracle.raygen_ao(rays, res, normals, 0, n_in, args.spp, AO_RADIUS, 0x9E3779B9)
    df, _, _ = oracle.raygen_ao(rays, res, normals, 0, n_in, args.spp, cam.far, 0x9E3779B9)
    hit_in = int((res[:n_in, 0] >= 0).sum())
    counted = W * H + 2 * hit_in * args.spp

    def step():
        trace(rays, True)
        trace(ao, False)
        trace(df, True)

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    sec = time.time() - t0
    value = counted * args.steps / sec * 1e-6
    sample = f"per step: {len(rays)} primary + {len(ao)} AO + {len(df)} diffuse rays (one <=1Mi batch of each secondary type), {what}, {threads} threads"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "conference stand-in room(283000, seed=2): primary + AO(32spp, r=5) + diffuse(32spp), 1024x768; CPU arm traces a bounded sample per step",
                      "primary_hits": int(hits)},
           "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
           "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--builder", default="hlbvh", choices=["hlbvh", "lbvh"],
                    help="GPU builder: hlbvh = Renderer default HLBVHParams{true, 4, 8, 0.001} (Renderer.cpp:201-209); lbvh = hlbvhBits 10")
    ap.add_argument("--hlbvh-bits", type=int, default=2,
                    help="HLBVHParams.hlbvhBits for --builder hlbvh (reference Renderer default 4; 2 = finer SAH top level: better tree, +1 ms build)")
    ap.add_argument("--collapse", type=int, default=1, choices=[0, 1],
                    help="leaf formation of the GPU builder: 1 = SAH-guided collapse (north_star pipeline, maxLeaf 8), 0 = the reference's count rule")
    ap.add_argument("--overlap", type=int, default=1, help="1 (default): the device-timed leg queues its launches on two kernel streams "
                                                            "(nt_set_deferred(2)); 0: one stream, every batch waits for the tail of the one before")
    ap.add_argument("--kernel", default=DEFAULT_KERNEL, help="trace kernel name (nt_set_kernel)")
    ap.add_argument("--partition", default="deal", choices=["deal", "slices"], help="N > 1: how a ray type's frame buffer is split over the GPUs (see rank_ranges)")
    ap.add_argument("--config3", type=int, default=1, help="N > 1 only: also strong-scale BASELINE.json configs[3] (10.5 M triangles, diffuse) with the NCCL broadcast timed")
    ap.add_argument("--reference-gpu", type=int, default=1, help="rank 0: also time the reference's own kernels recompiled for sm_100a (oracle/_ref), when present")
    ap.add_argument("--submission", default="frame", choices=["frame", "batches"],
                    help="frame (default): a rank's share of each ray type in one persistent launch (nt_trace_batches); batches: one launch per <= 1 Mi-ray batch")
    ap.add_argument("--raygen-order", type=int, default=1, choices=[0, 1],
                    help="slot order of the secondary rays (nt_raygen_set_order): 0 = the reference generator's order, 1 (default) = the same rays in "
                         "direction-coherent order inside tiles of <= 1024 slots (written by the generator itself, no extra pass)")
    ap.add_argument("--build-leg", type=int, default=1, help="N = 1: also report GPU build times (bench tree, LBVH, 10 M-triangle soup) in the line")
    ap.add_argument("--profile", action="store_true", help="dev: only the device-timed region (for runs under ncu); prints no JSON")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
