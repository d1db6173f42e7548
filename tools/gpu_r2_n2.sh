#!/bin/bash
# final build on a 2-GPU box: the driver's bench command at N=2 (weak), the strong-scaling job, the N=1 line for the per-GPU ratio, and the
# real-NCCL parity test at world 2
set -u
out=gpurun_out; mkdir -p $out; tag=${1:-r2n}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522"
$T bench.py --gpus 2 --steps 20 --warmup 5 > $out/${tag}_bench_n2.json 2>$out/${tag}_bench_n2.err
$T bench.py --gpus 2 --strong --steps 10 --warmup 3 > $out/${tag}_bench_strong_n2.json 2>$out/${tag}_bench_strong_n2.err
python bench.py --gpus 1 --no-e2e --no-cpu-baseline --no-api-mode --no-configs --no-fusion-mode > $out/${tag}_bench_n1_quick.json 2>/dev/null
(python -m pytest tests/test_gpu_dp_nccl.py -q -rs -W ignore -p no:cacheprovider 2>&1 | tail -5) > $out/${tag}_nccl_parity.log 2>&1
cat $out/${tag}_nccl_parity.log
python - <<PY
import json
for f in ("${tag}_bench_n1_quick", "${tag}_bench_n2", "${tag}_bench_strong_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["per_gpu_GBps"], d["ms_per_step"], (d.get("dp_check") or {}), d["clocks"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
