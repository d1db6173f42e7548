#!/bin/bash
# In-step sweep of the per-tensor launch shapes (LSQB200_TUNE) on the bench's own step; usage: tools/gpu_tilesweep.sh <tag> spec spec ...
out=gpurun_out; mkdir -p $out; tag=$1; shift
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fusion-mode --no-plan-mode --no-api-mode --no-configs --no-strong"
for i in ${ROUNDS:-1 2}; do
for t in "$@"; do
  LSQB200_TUNE="$t" timeout 120 $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-55s value %.1f  bwd %.1f  ms %.4f' % ('$t', d['value'], d['roofline']['achieved'], d['ms_per_step']))"
done; done | tee $out/${tag}_tilesweep.txt
