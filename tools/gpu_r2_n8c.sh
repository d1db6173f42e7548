#!/bin/bash
# final build on one 8-GPU box: the driver's bench command at N=8 and N=4 (weak) and the N=1 line for the per-GPU ratio
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2o}
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > $out/${tag}_bench_n$n.json 2>$out/${tag}_bench_n$n.err
done
python bench.py --gpus 1 --no-e2e --no-cpu-baseline --no-api-mode --no-configs --no-fusion-mode --no-plan-mode --no-strong > $out/${tag}_bench_n1_quick.json 2>/dev/null
python - <<PY
import json
for f in ("${tag}_bench_n1_quick", "${tag}_bench_n4", "${tag}_bench_n8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["per_gpu_GBps"], d["ms_per_step"], (d.get("dp_check") or {}).get("ok"), d["clocks"])
    except Exception as e:
        print(f, "FAILED", e)
PY
