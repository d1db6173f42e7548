#!/bin/bash
# usage: tools/ab_tune.sh tag "spec1" "spec2" ...   -> bench (no e2e) + bench_configs per LSQB200_TUNE spec
tag=$1; shift
i=0
for spec in "$@"; do
  LSQB200_TUNE="$spec" python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_$i.json 2>gpurun_out/${tag}_bench_$i.err
  LSQB200_TUNE="$spec" python tools/bench_configs.py --iters 20 > gpurun_out/${tag}_cfg_$i.json 2>gpurun_out/${tag}_cfg_$i.err
  echo "$i: $spec" >> gpurun_out/${tag}_specs.txt
  i=$((i+1))
done
