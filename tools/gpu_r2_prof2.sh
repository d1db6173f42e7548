#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
N="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1; rx=$2; skip=$3; cnt=$4; shift 4
    timeout 300 $N --kernel-name regex:$rx --launch-skip $skip --launch-count $cnt -o /tmp/$name "$@" > $out/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
    ncu -i /tmp/$name.ncu-rep --page source --csv > $out/$name.source.csv 2>/dev/null
    rm -f /tmp/$name.ncu-rep; }
cap r2_prof_stats4 lsq_rowstats 2 1 python tools/statsprof.py 4
cap r2_prof_stats4bf lsq_rowstats 2 1 python tools/statsprof.py 4 bf16
cap r2_prof_stats7 lsq_rowstats_ring 2 1 python tools/statsprof.py 7
ls -la $out | tail -12
