#!/bin/bash
# several LSQB200_TUNE specs on the headline step, round-robin: tools/ab_multi.sh <rounds> "<spec>" "<spec>" ...
set -u
R=$1; shift
Q="--no-e2e --no-cpu-baseline --no-fusion-mode --no-api-mode --no-configs --no-strong --steps 30"
for r in $(seq $R); do
  for spec in "$@"; do
    LSQB200_TUNE="$spec" python bench.py $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spec=[$spec]', 'value', d['value'], 'ms', d['ms_per_step'], 'bwd', d['roofline']['achieved'], 'plan', d['plan_mode']['value'])"
  done
done
