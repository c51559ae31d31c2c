"""Dev tool: condense `ncu -i <rep> --page raw --csv` into one markdown table row per kernel launch (duration, DRAM bytes and % of peak,
L1 / L2 hit rates, SM / memory throughput %, achieved occupancy, registers, lanes per instruction).  Usage:
    python scripts/ncu_summarise.py raw.csv [out.md]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr"), ("launch__grid_size", "grid")]


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], newline="") if not l.startswith("==")))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    out = ["| kernel | " + " | ".join(n for _, n in COLS) + " |", "|---|" + "---|" * len(COLS)]
    for r in body:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("nt::<unnamed>::", "").replace("<unnamed>::", "")
        cells = []
        for key, _ in COLS:
            if key not in idx:
                cells.append("-")
                continue
            v, u = r[idx[key]], units[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.1f}" if f < 1000 else f"{f:,.0f}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip())
        out.append(f"| `{name[:48]}` | " + " | ".join(cells) + " |")
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
