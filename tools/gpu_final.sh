#!/bin/bash
# Round-end GPU session: full GPU test suite (-rs: skips are visible), launch list + ncu --set full captures of the step's two dominant
# kernels (exported to CSV on the box: gpurun_out is capped at 64 MiB), host cost per public-op call, the bench line (with e2e and
# cpu_baseline) and the reference arm.  usage: tools/gpu_final.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-final}
out=gpurun_out
mkdir -p $out
(time python -m pytest tests -q -m gpu -rs -W ignore -p no:cacheprovider 2>&1 | tail -12) > $out/${tag}_all_tests.log 2>&1
Q="--no-e2e --no-cpu-baseline --no-plan-mode --no-fusion-mode --no-api-mode --no-configs --no-strong"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:lsq_ -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 $Q > $out/${tag}_launches.log 2>&1
cap() { name=$1; rx=$2; skip=$3; shift 3
    ncu --set full --clock-control none --import-source on --kernel-name regex:$rx --launch-skip $skip --launch-count 1 -f -o /tmp/$name "$@" > $out/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
    rm -f /tmp/$name.ncu-rep; }
cap ${tag}_prof_bwd lsq_flatbwd_kernel 69 python bench.py --steps 1 --warmup 3 $Q
cap ${tag}_prof_fwd lsq_flatfwd_kernel 1 python bench.py --steps 1 --warmup 3 $Q
python tools/host_overhead.py > $out/${tag}_host.txt 2>&1
python tools/host_overhead.py ref >> $out/${tag}_host.txt 2>&1
python bench.py > $out/${tag}_bench.json 2>$out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>$out/${tag}_bench_ref.err
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
tail -n 8 $out/${tag}_all_tests.log; tail -n 2 $out/${tag}_smoke.log; grep -v Warning $out/${tag}_host.txt | tail -3; cat $out/${tag}_bench_ref.json | cut -c1-300
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"])
print("api", {k: (v["ms_per_step"], v["vs_cabi_step"], v["host_us_per_site"]) for k, v in d["api_mode"].items() if isinstance(v, dict)})
print("plan", d["plan_mode"]["value"], "fusion", d["prologue_fusion"]["speedup"], "strong", d["strong_batch2048"]["value"])
for k, v in d["configs"].items():
    if isinstance(v, dict): print(" ", k, v["frac"], v["frac_stream"], v["ms"])
print(d["clocks"])
PY
