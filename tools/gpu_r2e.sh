#!/bin/bash
# Session r2e: column kernels - first loads ahead of the parameter round trip, merged channel runs in the backward epilogue, acq_rel fences.
set -u
out=gpurun_out; mkdir -p $out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_relu.py tests/test_gpu_add.py -q -x -W ignore -p no:cacheprovider -k "column or channel or config4" 2>&1 | tail -8) > $out/r2e_tests.log 2>&1
tail -3 $out/r2e_tests.log
for v in prev new noearly sfence; do
  L=$PWD/ab/lib_$v.so; [ $v = new ] && L=$PWD/lsqfakequantize-pytorch_b200/torchlsq/libtorchlsq_b200.so
  TORCHLSQ_B200_LIB=$L timeout 200 python tools/colbench3.py "pdl=1" 2>&1 | grep -v Warn | sed "s/^/$v /" > $out/r2e_colbench_$v.txt
done
cat $out/r2e_colbench_*.txt | grep "(256,2048,49)\|(50176,1024,1)\|(256,1024,196)" | sort -k2,3 -s
TORCHLSQ_B200_LIB=$PWD/ab/lib_trace.so timeout 200 python tools/coltrace.py "pdl=1" > $out/r2e_trace.txt 2>&1
echo ---- trace; cat $out/r2e_trace.txt | tail -12
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-fusion-mode --no-plan-mode --no-api-mode --no-configs --no-strong"
for i in 1 2; do for v in prev new sfence; do
  L=$PWD/ab/lib_$v.so; [ $v = new ] && L=$PWD/lsqfakequantize-pytorch_b200/torchlsq/libtorchlsq_b200.so
  TORCHLSQ_B200_LIB=$L timeout 200 $B > $out/r2e_bench_${v}_$i.json 2>$out/r2e_bench_${v}_$i.err
done; done
python - <<PY
import json,glob
for k in ("prev","new","sfence"):
    v=[json.loads(open(f).read().strip().splitlines()[-1]) for f in sorted(glob.glob("gpurun_out/r2e_bench_%s_*.json"%k))]
    print(k, [x["value"] for x in v], "bwd", [x["roofline"]["achieved"] for x in v])
PY
