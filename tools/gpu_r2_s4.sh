#!/bin/bash
# round 2, session 4: full GPU suite on the current build (C++ grouped op, tightened parity gates, full-size oracle tests) + full bench
set -u
out=gpurun_out; mkdir -p $out
(time python -m pytest tests -q -m gpu -x -rs 2>&1 | tail -15) > $out/r2s4_tests.log 2>&1
python bench.py --no-cpu-baseline > $out/r2s4_bench.json 2>$out/r2s4_bench.err
tail -n 15 $out/r2s4_tests.log; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s4_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"])
print(json.dumps(d["api_mode"], indent=0))
for k, v in d["configs"].items():
    if isinstance(v, dict): print(k, v["frac"], v["frac_stream"], v["ms"])
print(d["strong_batch2048"]["value"], d["e2e"]["value"])
PY
tail -3 $out/r2s4_bench.err; cat $out/parity_margins.json
