#!/bin/bash
# round 2: 2-GPU session - real-NCCL data-parallel parity test, the driver's bench command at N=2, and the strong-scaling job
set -u
out=gpurun_out; mkdir -p $out
(time python -m pytest tests/test_gpu_dp_nccl.py -q -m gpu -rs 2>&1 | tail -8) > $out/r2_n2_tests.log 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 20 --warmup 5 > $out/r2_bench_n2.json 2>$out/r2_bench_n2.err
$T bench.py --gpus 2 --strong --steps 10 --warmup 3 > $out/r2_bench_strong_n2.json 2>$out/r2_bench_strong_n2.err
tail -8 $out/r2_n2_tests.log; cat $out/r2_bench_n2.json; tail -3 $out/r2_bench_n2.err; cat $out/r2_bench_strong_n2.json; tail -3 $out/r2_bench_strong_n2.err
