#!/bin/bash
# Round-end GPU session: full GPU test suite, launch list + ncu --set full captures of the step's two dominant kernels, the bench line
# (with e2e and cpu_baseline) and the reference arm.  usage: tools/gpu_final.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-final}
out=gpurun_out
mkdir -p $out
(time python -m pytest tests -q -m gpu 2>&1 | tail -4) > $out/${tag}_all_tests.log 2>&1
Q="--no-e2e --no-cpu-baseline --no-plan-mode --no-fusion-mode"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:lsq_ -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 $Q > $out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:lsq_bwd_kernel --launch-skip 69 --launch-count 1 \
    -f -o $out/${tag}_prof_bwd python bench.py --steps 1 --warmup 3 $Q > $out/${tag}_prof_bwd.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:lsq_flatfwd_kernel --launch-skip 1 --launch-count 1 \
    -f -o $out/${tag}_prof_fwd python bench.py --steps 1 --warmup 3 $Q > $out/${tag}_prof_fwd.log 2>&1
python bench.py > $out/${tag}_bench.json 2>$out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>$out/${tag}_bench_ref.err
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
tail -n 4 $out/${tag}_all_tests.log $out/${tag}_smoke.log; cat $out/${tag}_bench.json
