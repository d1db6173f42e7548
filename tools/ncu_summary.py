#!/usr/bin/env python
"""Summarise Nsight Compute captures into the tracked profiles/ directory.

    python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep  profiles/r1_bwd_bf16.md [key=name ...]
    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md

`report` reads a --set full capture (ncu -i ... --page raw --csv) and writes the metrics the
roofline argument needs (duration, DRAM bytes, DRAM %, issue / pipe utilisation, occupancy,
registers); with key=name it also records dram bytes per launch in profiles/ncu_summary.json,
which bench.py reports as roofline.traffic.  `launches` aggregates a gpu__time_duration list.
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    return v * mult.get(u, 1)


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit.lower().replace("second", "s").replace("usecond", "us"), 1)


def report(rep, out_md, tags):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary: `{Path(rep).name}`", "",
             "Numbers under a profiler are not bench values: caches are flushed and kernels serialised; "
             "they show WHERE the time goes (DRAM bytes, pipe and issue utilisation).", ""]
    summary = {}
    for n, r in enumerate(rows):
        name = r[idx["Kernel Name"]]
        lines += [f"## launch {n}: `{name[:150]}`", "", "| metric | value |", "|---|---|"]
        for key, label in KEYS:
            if key in idx:
                lines.append(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        dur = to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
        lines += [f"| **DRAM traffic per launch** | {rd + wr:.0f} B ({(rd + wr) / 1e9:.4f} GB) |",
                  f"| **DRAM GB/s over the launch** | {(rd + wr) / dur / 1e3:.1f} |", ""]
        summary[n] = dict(kernel=name, duration_us=dur, dram_bytes_per_launch=rd + wr, dram_GBps=(rd + wr) / dur / 1e3)
    Path(out_md).write_text("\n".join(lines))
    if tags:
        js = Path(out_md).parent / "ncu_summary.json"
        data = json.loads(js.read_text()) if js.exists() else {}
        for t in tags:
            k, v = t.split("=")
            data[k] = dict(summary[int(v)], source=Path(rep).name, summary=Path(out_md).name)
        js.write_text(json.dumps(data, indent=1))
    print("wrote", out_md)


def launches(csv_path, out_md):
    """Per-kernel totals of a gpu__time_duration launch list.  Shares are taken over the PRODUCT kernels
    (namespace lsqb200) only; whatever else the process launched (torch RNG / fill kernels that build the
    synthetic inputs before the timed region) is listed separately."""
    lines = [ln for ln in open(csv_path) if not ln.startswith("==")]
    agg, other = OrderedDict(), OrderedDict()
    total = 0.0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        short = name.split("(")[0][:110]
        t = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        mine = "lsqb200::" in name or "lsq_" in name.split("(")[0]
        a = (agg if mine else other).setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += t
        if mine:
            total += t
    out = [f"# kernel launch list: `{Path(csv_path).name}`", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py`; per-launch times are cold-cache and "
           "serialised, so only each kernel's SHARE of the step is meaningful.", "",
           "## product kernels (the timed step)", "",
           "| kernel | launches | total us | share of step |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f} % |")
    if other:
        out += ["", "## other kernels in the capture (input generation and checks outside the timed region)", "",
                "| kernel | launches | total us |", "|---|---|---|"]
        for k, (n, t) in sorted(other.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{k}` | {n} | {t:.1f} |")
    Path(out_md).write_text("\n".join(out) + "\n")
    print("wrote", out_md)


if __name__ == "__main__":
    if sys.argv[1] == "report":
        report(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        launches(sys.argv[2], sys.argv[3])
