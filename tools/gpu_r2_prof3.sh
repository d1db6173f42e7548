#!/bin/bash
# round 2: ncu --set full of the export kernels; the quickstart example with and without grouping / fusion
set -u
out=gpurun_out; mkdir -p $out
name=r2_prof_export
timeout 300 ncu --set full --clock-control none --import-source on -f --kernel-name regex:lsq_ --launch-skip 4 --launch-count 2 -o /tmp/$name python tools/exportprof.py > $out/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null; rm -f /tmp/$name.ncu-rep
for flags in "" "--group-weights" "--group-weights --fuse-prologues"; do
  echo "== quickstart $flags"; timeout 300 python examples/qat_quickstart.py --steps 11 $flags 2>&1 | grep -v Warn | tail -6
done > $out/r2_quickstart.log 2>&1
cat $out/r2_quickstart.log
