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
# dram__bytes_read.sum + dram__bytes_write.sum per trace launch, from the ncu --set full captures of one primary, one AO
# and one diffuse batch weighted by the 1 + 24 + 24 launches of a step (profiles/r1_summary.md, round-1b captures: 35.1 / 37.0 / 53.9 MB)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 45.2e6
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ntrace_b200 import camera, capi, host

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host.init(local)
    dev = host.device()
    verts, tris, cam = make_workload()
    scene = host.Scene(verts, tris)

    # ---- BVH: GPU LBVH build on rank 0 (timed), NCCL broadcast of the three buffers to the replicas
    build_s = []
    capi.bvh_set_collapse(args.collapse, LEAF_SIZE)
    if rank == 0:
        for _ in range(1 + 5):
            build_s.append(capi.bvh_build(capi.BUILDER_HLBVH, scene.vtxPos, scene.triVtxIndex, scene.bboxMin, scene.bboxMax,
                                          args.hlbvh_bits if args.builder == "hlbvh" else 10, LEAF_SIZE, EPSILON))
    bcast_ms = 0.0
    if world > 1:
        from ntrace_b200 import multigpu
        bcast_ms = multigpu.broadcast_bvh(src=0) * 1e3
    (node_b, woop_b, idx_b), _ = capi.bvh_sizes()
    bvh = host.CudaBVH(layout=host.BVHLayout_Compact)
    bvh.resident = True
    tracer = host.CudaBVHTracer()
    tracer.setBVH(bvh)

    # ---- resident inputs: primary rays, then every AO / diffuse batch of the frame (weak scaling: every rank
    # traces a whole frame; ranks > 0 jitter their primaries so the rays differ)
    rg = host.RayGen(MAX_BATCH)
    prim = host.RayBuffer()
    rg.primary(prim, cam.position, camera.nscreen_to_world(cam, W, H), W, H, cam.far, 0 if rank == 0 else 1000 + rank)
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
    res_dev = torch.empty((MAX_BATCH, 4), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    traced = {k: sum(b[2] for b in batches if b[0] == k) for k in ("primary", "AO", "diffuse")}
    counted = {"primary": W * H, "AO": hits * args.spp, "diffuse": hits * args.spp}
    counted_step = sum(counted.values())
    ray_bytes = sum(b[2] for b in batches) * 32

    res_alt = [res_dev, torch.empty_like(res_dev)] if args.overlap else [res_dev, res_dev]

    def step():
        for i, (_, rays, n, closest) in enumerate(batches):
            capi.trace_batch(rays, res_alt[i & 1], n, closest)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        capi.synchronize()

    # ---- device-timed region: W warm-up steps, then exactly K steps between two events on the launching stream
    capi.set_deferred(2 if args.overlap else 1)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local, gpu_uuid(torch, local))
    sampler.start()
    l0 = capi.launch_count()
    capi.event_record(0)
    for _ in range(args.steps):
        step()
    capi.event_record(1)
    sec = capi.event_elapsed(0, 1)
    launches = capi.launch_count() - l0
    barrier()
    clocks = sampler.stop()
    capi.set_deferred(False)

    if args.profile:
        if rank == 0:
            print(f"profile run: {counted_step * args.steps / sec * 1e-6:.1f} Mrays/s (number taken under a profiler is not a bench value)")
        return None

    # per-type kernel time (synchronous calls, CUDA events around each launch: the reference's accounting)
    type_sec = {"primary": 0.0, "AO": 0.0, "diffuse": 0.0}
    for name, rays, n, closest in batches:
        type_sec[name] += capi.trace_batch(rays, res_dev, n, closest)

    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), H2D of the rays and D2H of the
    # results inside the timed region, every step
    from ntrace_b200 import multigpu
    prev_affinity, numa_node = multigpu.bind_host_to_gpu(local)      # pinned buffers land on the GPU's NUMA node
    host_batches = []
    for name, rays, n, closest in batches:
        hb = torch.empty((n, 8), dtype=torch.float32, pin_memory=True)
        hb.copy_(rays)
        host_batches.append((hb, n, closest))
    res_host = torch.empty((MAX_BATCH, 4), dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    for hb, n, closest in host_batches:          # one warm-up pass (staging buffers grow here)
        capi.trace_batch(hb, res_host, n, closest)
    barrier()
    capi.event_record(2)
    for _ in range(e2e_steps):
        for hb, n, closest in host_batches:
            capi.trace_batch(hb, res_host, n, closest)
    capi.event_record(3)
    e2e_sync_sec = capi.event_elapsed(2, 3)
    barrier()
    # pipelined form: nt_trace_batch_async keeps NSLOT independent batches of the frame in flight (H2D of batch i+1,
    # traversal of batch i and D2H of batch i-1 overlap); every batch's rays still cross PCIe in and its results out
    NSLOT = 3
    res_slots = [torch.empty((MAX_BATCH, 4), dtype=torch.int32, pin_memory=True) for _ in range(NSLOT)]

    def pipelined_frame():
        for i, (hb, n, closest) in enumerate(host_batches):
            s = i % NSLOT
            capi.trace_wait(s)
            capi.trace_batch_async(hb, res_slots[s], n, closest, s)
        for s in range(NSLOT):
            capi.trace_wait(s)

    pipelined_frame()                            # warm-up (slot staging grows here)
    barrier()
    capi.event_record(2)
    for _ in range(e2e_steps):
        pipelined_frame()
    capi.event_record(3)
    e2e_sec = capi.event_elapsed(2, 3)
    barrier()
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)                        # the CPU baseline leg uses every core again

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([sec, e2e_sec, e2e_sync_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec, e2e_sec, e2e_sync_sec = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([counted_step, launches], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        counted_all, launches_all = float(c[0]), int(c[1])
    else:
        counted_all, launches_all = float(counted_step), launches

    value = counted_all * args.steps / sec * 1e-6
    e2e_value = counted_all * e2e_steps / e2e_sec * 1e-6

    out = None
    if rank == 0:
        peak, peak_src = load_peaks()
        # strided ~500K-ray sample of every ray type, taken across all batches of the frame
        samples = []
        for k, closest in (("primary", True), ("AO", False), ("diffuse", True)):
            stride = max(1, traced[k] // 500_000)
            samples.append((k, torch.cat([b[1][::stride] for b in batches if b[0] == k]).cpu().numpy(), closest))
        cpu = cpu_baseline_leg(verts, tris, cam, args, gpu_bvh=capi.bvh_download()[:3], sample_batches=samples)
        ref_gpu_rows = reference_gpu_leg(torch, capi, batches)
        # algorithmic bytes of one step = sum over ray types of mean B_ray (oracle counters on the
        # GPU-built BVH, ~500K-ray strided sample per type) x rays traced of that type
        alg_bytes_step = sum(cpu["bytes_per_ray"][k] * traced[k] for k in traced)
        n_launch_step = len(batches)
        avg_launch_s = sec / (args.steps * n_launch_step)
        achieved = alg_bytes_step / n_launch_step / avg_launch_s * 1e-9
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "conference stand-in room(283000, seed=2): primary + AO(32spp, r=5, any-hit) + diffuse(32spp, closest-hit), "
                                   "1024x768, <=1Mi rays/batch, GPU %s leaf 8%s" % ("HLBVH(hlbvhBits %d)" % args.hlbvh_bits if args.builder == "hlbvh" else "LBVH", ", SAH-guided collapse" if args.collapse else ""),
                       "rays_traced_per_step_per_gpu": int(sum(traced.values())), "rays_counted_per_step_per_gpu": int(counted_step),
                       "batches_per_step": n_launch_step, "kernel": "b200_persistent_speculative_while_while",
                       "submission": ("nt_set_deferred(2): the step's launches are queued on two kernel streams, consecutive batches overlap at their tails"
                                      if args.overlap else "nt_set_deferred(1): the step's launches are queued on one stream"),
                       "l2": "inputs exceed L2: %.0f MB of rays per step stream from HBM; the %.0f MB BVH is reused within a frame by design"
                             % (ray_bytes / 1e6, (node_b + woop_b + idx_b) / 1e6),
                       "parallelism": "ray batches sharded per GPU, BVH replicated by NCCL broadcast" if world > 1 else "single GPU"},
            "detail": {"primary_mrays": counted["primary"] / type_sec["primary"] * 1e-6, "ao_mrays": counted["AO"] / type_sec["AO"] * 1e-6,
                       "diffuse_mrays": counted["diffuse"] / type_sec["diffuse"] * 1e-6,
                       "build_ms": float(np.mean(build_s[1:]) * 1e3), "build_mtris": len(tris) / float(np.mean(build_s[1:])) * 1e-6,
                       "bvh_broadcast_ms": bcast_ms, "primary_hits": int(hits)},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(ray_bytes), "d2h_bytes_per_step": int(ray_bytes // 2),
                    "steps": e2e_steps, "host_numa_node": numa_node,
                    "api": "nt_trace_batch_async/nt_trace_wait, 3 batches in flight, pinned host rays in and results out every batch",
                    "sync_value": counted_all * e2e_steps / e2e_sync_sec * 1e-6,
                    "sync_api": "nt_trace_batch (one synchronous call per batch, zero-copy pinned buffers)"},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes (nodes+triangles fetched per ray, oracle-counted) over avg launch time (timed region / launches: with the overlapped submission consecutive launches share the GPU at their tails); the BVH is L2-resident, "
                                 "so DRAM traffic is far below the algorithmic bytes and frac can exceed 1 (see profiles/)",
                         "bytes_per_ray": cpu["bytes_per_ray"],
                         # what actually limits the kernel, from the committed ncu captures of this workload (static numbers,
                         # profiles/r1b_prof2_trace_s{50,52,80}_raw.csv): issue slots busy, useful lanes per instruction, DRAM throughput
                         "ncu": {"issue_active_pct": {"primary": 66.5, "AO": 71.6, "diffuse": 59.6},
                                 "lanes_per_instruction": {"primary": 21.3, "AO": 12.3, "diffuse": 12.5},
                                 "dram_throughput_pct_of_peak": 2.0, "l1tex_throughput_pct": {"primary": 51.5, "AO": 67.6, "diffuse": 75.0},
                                 "limiter": "instruction issue x SIMD efficiency, then L1/L2 latency; not HBM (see profiles/r1_summary.md)"}},
            "cpu_baseline": cpu["baseline"],
            "reference_gpu": ref_gpu_rows,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# --------------------------------------------------------------------------------------------------
def cpu_baseline_leg(verts, tris, cam, args, gpu_bvh=None, sample_batches=None):
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
    if gpu_bvh is not None:
        for name, r, closest in sample_batches:
            res, cnt = oracle.compact_trace(gpu_bvh[0], gpu_bvh[1], gpu_bvh[2], r, closest, counters=True, nthreads=threads)
            bpr[name] = bytes_per_ray(cnt, float((res[:, 0] >= 0).mean()))
    baseline = {"value": total_rays / total_s * 1e-6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                "sample": "restated BVH::trace (BVH.cpp:90-186) on a restated SplitBVH (alpha 1e-5, leaf 1/1) of the same scene, OpenMP over rays, "
                          + ", ".join(f"{len(r)} {n}" for n, r, _ in sample_batches) + " rays",
                "build_s_1thread": build_s, "build_mtris_1thread": len(tris) / build_s * 1e-6, "splitbvh_sah": st.sah}
    return {"baseline": baseline, "bytes_per_ray": bpr}


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
    ap.add_argument("--profile", action="store_true", help="dev: only the device-timed region (for runs under ncu); prints no JSON")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
