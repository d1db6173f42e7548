#!/bin/bash
# round 2, session 1: the C++ binding on the GPU - full GPU suite, host cost per call, bench N=1
set -u
out=gpurun_out; mkdir -p $out
(time python -m pytest tests -q -m gpu -x -rs 2>&1 | tail -15) > $out/r2s1_tests.log 2>&1
python tools/host_overhead.py > $out/r2s1_host.txt 2>&1
python tools/host_overhead.py ref >> $out/r2s1_host.txt 2>&1
python bench.py --no-cpu-baseline > $out/r2s1_bench.json 2>$out/r2s1_bench.err
tail -n 15 $out/r2s1_tests.log; cat $out/r2s1_host.txt; cat $out/r2s1_bench.json; tail -5 $out/r2s1_bench.err
