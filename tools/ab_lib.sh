#!/bin/bash
# A/B of two library builds on one box: ab/lib_prev.so against the in-tree library, alternating, bench.py without the side legs.
# usage: tools/ab_lib.sh <tag> [rounds]
tag=${1:-ab}; rounds=${2:-3}
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fusion-mode --no-plan-mode"
mkdir -p gpurun_out
for i in $(seq 1 $rounds); do
  TORCHLSQ_B200_LIB=$PWD/ab/lib_prev.so $B > gpurun_out/${tag}_prev_$i.json 2>gpurun_out/${tag}_prev_$i.err
  $B > gpurun_out/${tag}_new_$i.json 2>gpurun_out/${tag}_new_$i.err
done
python - <<PY
import json,glob
for k in ("prev","new"):
    v=[json.load(open(f)) for f in sorted(glob.glob("gpurun_out/${tag}_%s_*.json"%k))]
    print(k, [x["value"] for x in v], "bwd", [x["roofline"]["achieved"] for x in v])
PY
