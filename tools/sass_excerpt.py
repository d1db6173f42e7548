#!/usr/bin/env python
"""Opcode counts per hot kernel from `cuobjdump -sass` of the shipped library -> profiles/r2_sass_excerpt.md
(evidence for the 256-bit global accesses, the bulk copies of the staged kernels, and the absence of tensor-core opcodes).
    python tools/sass_excerpt.py [out.md]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "lsqfakequantize-pytorch_b200" / "torchlsq" / "libtorchlsq_b200.so"
WANT = ['lsq_flatfwd_kernel<__nv_bfloat16', 'lsq_bwd_kernel<__nv_bfloat16, 0, 8, 0, 256', 'lsq_flatbwd_kernel<__nv_bfloat16, 0, 0',
        'lsq_fwd_kernel<float, 0, 8, false, 256', 'lsq_bwd_kernel<float, 0, 8, 0, 256', 'lsq_rowfwd_kernel<float', 'lsq_rowbwd_kernel<float, 0, 0',
        'lsq_rowstats3_kernel<float, 256, 2', 'lsq_rowstats_ring_kernel<float', 'lsq_observe_kernel<__nv_bfloat16, 8, 256',
        'lsq_col_fwd_kernel<__half, 0, false, 4, 4', 'lsq_col_bwd_kernel<__half, 0, 0, 4, 4', 'lsq_col_bwd_tma_kernel<__half, 0, 0, 4, 3',
        'lsq_stats_kernel<float, 8, 256', 'lsq_plan_patch_kernel']


def main():
    out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r2_sass_excerpt.md"
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    rows = []
    for f in re.split(r'\n\s*Function : ', txt)[1:]:
        name = f.split('\n', 1)[0].strip()
        dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
        c = collections.Counter(m.group(1) for m in re.finditer(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', f, re.M))
        rows.append((dem, c, sum(c.values())))
    pick = lambda c, pat: sum(v for k, v in c.items() if re.search(pat, k))
    out = ["# SASS evidence: wide loads / stores and bulk copies per hot kernel (round 2)", "",
           "`cuobjdump -sass torchlsq/libtorchlsq_b200.so` (sm_100a, the shipped build), opcode counts per kernel by `tools/sass_excerpt.py`.",
           "256-bit global accesses (`LDG.E...256` / `STG.E...256`, new on sm_100) carry every row-tiled / flat / weight-row kernel; the column kernels use",
           "128-bit units (a unit must not straddle more channels than a thread holds constants for); the bulk-copy kernels show `UBLKCP` and `SYNCS` (mbarrier).",
           "No tensor-core opcodes (`UTC*MMA`, `HMMA`) anywhere: the path is not a contraction.", "",
           "| kernel (first matching instantiation) | SASS instr | LDG .256 | LDG .128 | LDG other | STG .256 | STG .128 | UBLKCP | SYNCS (mbarrier) | FFMA | FRND | FMNMX | DADD/DFMA/DMUL | RED/ATOM |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for w in WANT:
        m = [r for r in rows if w in r[0]]
        if not m:
            out.append(f"| `{w}` | not found |")
            continue
        dem, c, tot = m[0]
        l256, l128, lall = pick(c, r'^LDG\..*256'), pick(c, r'^LDG\..*128'), pick(c, r'^LDG')
        cut = dem.find('(', max(dem.find('kernel'), 0))
        out.append(f"| `{(dem[:cut] if cut > 0 else dem)[:95]}` | {tot} | {l256} | {l128} | {lall - l256 - l128} | {pick(c, r'^STG\..*256')} | {pick(c, r'^STG\..*128')} | "
                   f"{pick(c, r'^UBLKCP')} | {pick(c, r'^SYNCS')} | {pick(c, r'^FFMA')} | {pick(c, r'^FRND')} | {pick(c, r'^FMNMX')} | "
                   f"{pick(c, r'^D(ADD|FMA|MUL)')} | {pick(c, r'^(RED|ATOM)')} |")
    mma = sum(pick(c, r'MMA') for _, c, _ in rows)
    out += ["", f"Kernels in the library: {len(rows)}; opcodes matching `MMA` over all of them: {mma}."]
    out_path.write_text("\n".join(out) + "\n")
    print(out_path)


if __name__ == "__main__":
    main()
