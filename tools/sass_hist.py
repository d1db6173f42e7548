#!/usr/bin/env python
"""Histogram of executed SASS opcodes (and stall samples) of one kernel in an ncu report:
    python tools/sass_hist.py report.ncu-rep <kernel-substring> [elements]
Reads `ncu --page source --print-source sass --csv`; prints warp instructions per opcode, and per element when
`elements` is given."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, needle = sys.argv[1], sys.argv[2]
elements = float(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for b in blocks:
    name = next(csv.reader([b[0]]))[1]
    if needle not in name:
        continue
    rows = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
    ops, stalls = Counter(), Counter()
    total = 0
    for r in rows:
        src = r["Source"].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0] if src else "?"
        full = ".".join(src.split()[0].split(".")[:3])
        n = int(r["Instructions Executed"] or 0)
        ops[full] += n
        stalls[full] += int(r["# Samples"] or 0)
        total += n
    print(name[:120])
    print("total warp instructions", total, "" if not elements else f"= {32 * total / elements:.2f} thread-instr / element")
    for op, n in ops.most_common(40):
        per = f"{32 * n / elements:6.2f}/el" if elements else ""
        print(f"  {op:28s} {n:12d} {100 * n / total:5.1f}%  {per}  samples {stalls[op]}")
    break
