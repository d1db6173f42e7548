#!/bin/bash
# round 2, final build: the driver's bench command at N=2, 4, 8 (weak) and the strong-scaling job at N=8, on one 8-GPU box
set -u
out=gpurun_out; mkdir -p $out
for n in 2 4 8; do
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n"
  $T bench.py --gpus $n --steps 20 --warmup 5 > $out/r2c_bench_n$n.json 2>$out/r2c_bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --strong --steps 10 --warmup 3 > $out/r2c_bench_strong_n8.json 2>$out/r2c_bench_strong_n8.err
python bench.py --gpus 1 --no-e2e --no-cpu-baseline --no-api-mode --no-configs --no-fusion-mode > $out/r2c_bench_n1_quick.json 2>/dev/null
python - <<'PY'
import json
for f in ("r2c_bench_n1_quick", "r2c_bench_n2", "r2c_bench_n4", "r2c_bench_n8", "r2c_bench_strong_n8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["per_gpu_GBps"], d["ms_per_step"], (d.get("dp_check") or {}).get("ok"), d["clocks"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
