"""Measure what binds the trace kernel ON THE BINARY THAT IS BEING TIMED and write it where bench.py reads it.

Runs on the GPU box (under gpurun, one GPU):
    python scripts/ncu_binding.py --kernel b200_wide4 --out profiles/r2_binding_b200_wide4.json
It (1) measures an L2 bandwidth peak (torch copy of an L2-resident buffer, CUDA events) next to MEASURED_PEAKS.json's HBM peak,
(2) runs `ncu --metrics ...` over scripts/profile_kernel.py (one primary, one AO, one diffuse launch of the chosen kernel on the bench
frame) and (3) stores per ray type: duration, lanes per instruction, issue-slot utilisation, ALU / FMA / LSU pipe utilisation,
L1 data-pipe (wavefront) utilisation, L1 / L2 hit rates, L2 bytes and GB/s against the measured L2 peak, DRAM bytes and GB/s
against the measured HBM peak, the stall breakdown, plus the sha256 of the library's sources (ntrace_b200.build.source_sha16) so that bench.py can tell whether the
capture describes the library it is timing.  Numbers taken under ncu are NOT bench values: durations here are cold-cache and
serialised; bench.py uses the RATES and FRACTIONS only.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def lib_sha():
    from ntrace_b200 import build as nb
    return nb.source_sha16()


def measure_l2_peak():
    """GB/s (read + write bytes) of a device copy whose 2 x 24 MB working set stays in the 126 MB L2; best of 20, CUDA events."""
    import torch
    n = 24 << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    best = 0.0
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 8 * 2 * n / (e0.elapsed_time(e1) * 1e-3) * 1e-9)
    return best


def parse_ncu_csv(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per_launch = {}
    for r in rd:
        lid = r.get("ID")
        if lid is None:
            continue
        d = per_launch.setdefault(lid, {"kernel": r.get("Kernel Name", "")})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            pass
        d.setdefault("_units", {})[r.get("Metric Name")] = r.get("Metric Unit")
    for lid in sorted(per_launch, key=lambda x: int(x)):
        rows.append(per_launch[lid])
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="b200_persistent_speculative_while_while")
    ap.add_argument("--scene", default="conference")
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--out", required=True)
    ap.add_argument("--regex", default="trace_")
    ap.add_argument("--hlbvh-bits", type=int, default=2)
    ap.add_argument("--raygen-order", type=int, default=1, help="slot order of the secondary rays (nt_raygen_set_order), as bench.py's default")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    raw = os.path.splitext(args.out)[0] + "_ncu.csv"
    l2_peak = measure_l2_peak()
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", f"regex:{args.regex}", "--csv", "--log-file", raw,
           sys.executable, os.path.join(ROOT, "scripts", "profile_kernel.py"), "--kernel", args.kernel, "--scene", args.scene, "--batch", str(args.batch), "--hlbvh-bits", str(args.hlbvh_bits), "--raygen-order", str(args.raygen_order)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    sys.stdout.write(r.stdout[-2000:])
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise SystemExit(r.returncode)
    rows = parse_ncu_csv(raw)
    if len(rows) < 4:
        raise SystemExit(f"expected >= 4 profiled launches, got {len(rows)}")
    rows = rows[-3:]                                   # primary, AO, diffuse (the launch before them is the warm-up)
    hbm_peak = 6650.0
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"kernel": args.kernel, "scene": args.scene, "raygen_order": args.raygen_order, "lib_sha16": lib_sha(), "l2_peak_gbs_measured": l2_peak, "hbm_peak_gbs": hbm_peak,
           "how": "ncu --metrics (list in scripts/ncu_binding.py) --clock-control none over scripts/profile_kernel.py: one launch per ray type on the bench frame; "
                  "L2 peak = torch copy of a 2 x 24 MB L2-resident working set, best of 20 (read + write bytes)",
           "per_type": {}}
    for name, d in zip(("primary", "AO", "diffuse"), rows):
        units = d.pop("_units", {})
        ns = d["gpu__time_duration.sum"] * (1e3 if units.get("gpu__time_duration.sum") in ("us", "usecond") else 1e6 if units.get("gpu__time_duration.sum") in ("ms", "msecond") else 1.0)
        sec = ns * 1e-9
        scale = lambda k: d[k] * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units.get(k), 1.0)
        l2b, drr, drw = scale("lts__t_bytes.sum"), scale("dram__bytes_read.sum"), scale("dram__bytes_write.sum")
        fr = {
            "issue_slots": d["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100,
            "alu_pipe": d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"] / 100,
            "fma_pipe": d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"] / 100,
            "lsu_pipe": d["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"] / 100,
            "l1_data_pipe": d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"] / 100,
            "l2_bandwidth": l2b / sec * 1e-9 / l2_peak,
            "hbm_bandwidth": (drr + drw) / sec * 1e-9 / hbm_peak,
        }
        top = max(fr, key=fr.get)
        out["per_type"][name] = {
            "kernel_symbol": d["kernel"], "duration_us_under_ncu": sec * 1e6,
            "lanes_per_instruction": d["smsp__thread_inst_executed_per_inst_executed.ratio"],
            "warp_instructions": d["smsp__inst_executed.sum"],
            "fractions_of_peak": fr, "binding": top, "binding_frac": fr[top],
            "l1_hit_rate": d["l1tex__t_sector_hit_rate.pct"] / 100, "l2_hit_rate": d["lts__t_sector_hit_rate.pct"] / 100,
            "l1_sectors_global_ld": d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"], "l1_requests_global_ld": d["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"],
            "l2_bytes": l2b, "l2_gbs": l2b / sec * 1e-9, "dram_bytes": drr + drw, "dram_gbs": (drr + drw) / sec * 1e-9,
            "warps_active_frac": d["sm__warps_active.avg.pct_of_peak_sustained_active"] / 100,
            "registers_per_thread": d.get("launch__registers_per_thread"), "grid": d.get("launch__grid_size"),
            "stall_warps_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: v for k, v in d.items() if "issue_stalled_" in k},
        }
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: {"binding": v["binding"], "frac": round(v["binding_frac"], 3), "lanes": round(v["lanes_per_instruction"], 2),
                          "us": round(v["duration_us_under_ncu"], 1), "fr": {a: round(b, 3) for a, b in v["fractions_of_peak"].items()}} for k, v in out["per_type"].items()}))


if __name__ == "__main__":
    main()
