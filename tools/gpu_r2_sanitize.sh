#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over tools/sanitize_smoke.py - every kernel family, round-2 paths included
set -u
out=gpurun_out; mkdir -p $out
: > $out/r2_compute_sanitizer.txt
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool" >> $out/r2_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py 2>&1 | grep -E "sanitize_smoke ok|ERROR SUMMARY|RACECHECK SUMMARY|=========.*(Invalid|Race|Uninit|hazard|Barrier|error)" | head -20 >> $out/r2_compute_sanitizer.txt
done
cat $out/r2_compute_sanitizer.txt
