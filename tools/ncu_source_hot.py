#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` export: per kernel, total samples by stall reason, the instruction mix by
executed-count class (setup = executed once per warp, loop = more) and the top-N SASS lines by stall samples.
    python tools/ncu_source_hot.py gpurun_out/x.source.csv [kernel-substring] [topN]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for k in kernels:
        if want not in k["name"]:
            continue
        h = {n: i for i, n in enumerate(k["hdr"])}
        stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        tot = defaultdict(int)
        samples = 0
        insts = 0
        minexec = min((int(r[h["Instructions Executed"]]) for r in k["rows"] if int(r[h["Instructions Executed"]]) > 0), default=1)
        setup_inst = loop_inst = 0
        for r in k["rows"]:
            s = int(r[h["# Samples"]])
            samples += s
            e = int(r[h["Instructions Executed"]])
            insts += e
            if e <= 1.01 * max(minexec, 1) * 1.0 or e <= int(k["rows"][0][h["Instructions Executed"]]):
                setup_inst += e
            else:
                loop_inst += e
            for c in stall_cols:
                tot[c] += int(r[h[c]])
        print(f"\n## {k['name'][:140]}\nSASS lines {len(k['rows'])}, warp instructions {insts} (executed <= once per warp: {setup_inst}, more: {loop_inst}), samples {samples}")
        print("stalls: " + ", ".join(f"{c[6:]}={v} ({100*v/max(samples,1):.0f}%)" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > samples))
        hot = sorted(k["rows"], key=lambda r: -int(r[h["# Samples"]]))[:top]
        for r in hot:
            reasons = sorted(((int(r[h[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
            print(f"  {int(r[h['# Samples']]):6d}  exec={r[h['Instructions Executed']]:>8}  {r[h['Source']].strip()[:70]:70s} {reasons}")


if __name__ == "__main__":
    main()
