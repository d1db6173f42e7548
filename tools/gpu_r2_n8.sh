#!/bin/bash
# round 2: 8-GPU session - NCCL parity at world 4, the driver's bench command at N=4 and N=8, strong scaling at N=4 / N=8
set -u
out=gpurun_out; mkdir -p $out
(time python -m pytest tests/test_gpu_dp_nccl.py -q -m gpu -rs 2>&1 | tail -6) > $out/r2_n8_tests.log 2>&1
for n in 4 8; do
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"
  $T bench.py --gpus $n --steps 20 --warmup 5 > $out/r2_bench_n$n.json 2>$out/r2_bench_n$n.err
  $T bench.py --gpus $n --strong --steps 10 --warmup 3 > $out/r2_bench_strong_n$n.json 2>$out/r2_bench_strong_n$n.err
done
tail -6 $out/r2_n8_tests.log
for n in 4 8; do python - <<PY
import json
for f in ("gpurun_out/r2_bench_n$n.json", "gpurun_out/r2_bench_strong_n$n.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["per_gpu_GBps"], d["ms_per_step"], d["dp_check"], d["clocks"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -2 $out/r2_bench_n$n.err
done
