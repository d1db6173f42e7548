#!/bin/bash
# ncu --set full of the column kernels (fwd + bwd, one launch each) on the three short-row layouts; raw metrics exported to CSV on the box
set -u
out=gpurun_out; mkdir -p $out
cap() { name=$1; shift
    timeout 170 ncu --set full --clock-control none -f --kernel-name regex:lsq_col --launch-skip 4 --launch-count 2 -o /tmp/$name "$@" > $out/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null; rm -f /tmp/$name.ncu-rep; }
cap r2q_prof_col7 python tools/colprof.py 256 2048 49
cap r2q_prof_colcl python tools/colprof.py 50176 1024 1
cap r2q_prof_col14 python tools/colprof.py 256 1024 196
python tools/ncu_csv_summary.py $out/r2q_prof_col7.raw.csv $out/r2q_prof_colcl.raw.csv $out/r2q_prof_col14.raw.csv
