#!/bin/bash
# In-step A/B of library builds (ab/lib_<tag>.so against the in-tree library) x tuning specs; usage: tools/gpu_libsweep.sh <out-tag> "<lib tags>" spec ...
out=gpurun_out; mkdir -p $out; tag=$1; libs=$2; shift 2
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fusion-mode --no-plan-mode --no-api-mode --no-configs --no-strong"
for i in 1 2; do for l in $libs; do for t in "$@"; do
  L=$PWD/ab/lib_$l.so; [ $l = new ] && L=$PWD/lsqfakequantize-pytorch_b200/torchlsq/libtorchlsq_b200.so
  TORCHLSQ_B200_LIB=$L LSQB200_TUNE="$t" timeout 120 $B 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-8s %-40s value %.1f  bwd %.1f  ms %.4f' % ('$l', '$t', d['value'], d['roofline']['achieved'], d['ms_per_step']))"
done; done; done | tee $out/${tag}_libsweep.txt
