#!/bin/bash
# round 2, session 2: grouped weight quantizers + the full bench line (configs, strong, api_mode)
set -u
out=gpurun_out; mkdir -p $out
(time python -m pytest tests/test_gpu_group.py tests/test_gpu_ops.py -q -m gpu -x -rs 2>&1 | tail -25) > $out/r2s2_tests.log 2>&1
python bench.py --no-cpu-baseline > $out/r2s2_bench.json 2>$out/r2s2_bench.err
tail -n 25 $out/r2s2_tests.log; cat $out/r2s2_bench.json; tail -5 $out/r2s2_bench.err
