#!/bin/bash
# round 2: ncu --set full captures of the non-dominant kernel families (column, weight rows, statistics, observer).
# The reports are exported to CSV on the box (raw metrics + per-instruction source page) and deleted: gpurun_out is capped at 64 MiB.
set -u
out=gpurun_out; mkdir -p $out
N="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, count, command...
    name=$1; rx=$2; skip=$3; cnt=$4; shift 4
    $N --kernel-name regex:$rx --launch-skip $skip --launch-count $cnt -o /tmp/$name "$@" > $out/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
    ncu -i /tmp/$name.ncu-rep --page source --csv > $out/$name.source.csv 2>/dev/null
    rm -f /tmp/$name.ncu-rep
}
cap r2_prof_col7 lsq_col 4 2 python tools/colprof.py 256 2048 49
cap r2_prof_colcl lsq_col 4 2 python tools/colprof.py 50176 1024 1
cap r2_prof_col14 lsq_col 4 2 python tools/colprof.py 256 1024 196
cap r2_prof_rows lsq_row 4 4 python tools/planprof.py
cap r2_prof_obs lsq_observe 2 1 python tools/obsprof.py
du -sh $out; ls -la $out | tail -20
