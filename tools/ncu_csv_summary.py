#!/usr/bin/env python
"""Summarise `ncu -i rep --page raw --csv` exports (one row per profiled launch) as a markdown table.
    python tools/ncu_csv_summary.py gpurun_out/r2_prof_col7.raw.csv [more.csv ...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration us", 1e-3), ("dram__bytes_read.sum", "DRAM read MB", 1e-6), ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1), ("smsp__inst_executed.sum", "warp instr", 1),
        ("launch__occupancy_limit_registers", "occ lim regs", 1), ("launch__occupancy_limit_shared_mem", "occ lim smem", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_sb", 1)]
UNIT = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"\n### {path}")
        for r in data:
            name = r[idx["Kernel Name"]]
            out = [name[:110]]
            for key, label, _ in KEYS:
                if key not in idx:
                    continue
                v, u = r[idx[key]].replace(",", ""), units[idx[key]]
                try:
                    f = float(v)
                except ValueError:
                    continue
                if u in UNIT:
                    f *= UNIT[u]
                out.append(f"{label}={f:.4g}")
            print("- " + "; ".join(out))


if __name__ == "__main__":
    main()
