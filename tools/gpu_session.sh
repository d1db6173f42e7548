#!/bin/bash
# One GPU-box session: A/B of two library builds, per-config numbers, launch list, ncu --set full captures.
# usage: tools/gpu_session.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-s}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fusion-mode"
if [ -f ab/lib_prev.so ]; then
  for i in 1 2; do
    TORCHLSQ_B200_LIB=$PWD/ab/lib_prev.so $B > $out/${tag}_ab_prev_$i.json 2>$out/${tag}_ab_prev_$i.err
    $B > $out/${tag}_ab_new_$i.json 2>$out/${tag}_ab_new_$i.err
  done
fi
python tools/bench_configs.py > $out/${tag}_configs.json 2>$out/${tag}_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:lsq_ -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-plan-mode --no-fusion-mode > $out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:lsq_bwd_kernel --launch-skip 69 --launch-count 2 \
    -f -o $out/${tag}_prof_bwd python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-plan-mode --no-fusion-mode > $out/${tag}_prof_bwd.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:lsq_flatfwd_kernel --launch-skip 1 --launch-count 1 \
    -f -o $out/${tag}_prof_fwd python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-plan-mode --no-fusion-mode > $out/${tag}_prof_fwd.log 2>&1
python bench.py > $out/${tag}_bench.json 2>$out/${tag}_bench.err
tail -n 3 $out/${tag}_ab_*.json $out/${tag}_bench.json
