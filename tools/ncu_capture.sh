#!/bin/bash
# One `ncu --set full` capture of the solve kernel of a bench workload. usage: ncu_capture.sh OUTDIR TAG workload [env...]
# (run only after the same bench command has exited 0 without ncu; numbers printed under ncu are not bench values).
# The raw-metrics page is exported to OUTDIR/raw_TAG.csv on the box; the .ncu-rep itself (~15 MB, gpurun_out/ is capped
# at 64 MiB per call) is kept only with KEEP_REP=1.
out=$1; tag=$2; wl=$3; shift 3
mkdir -p $out
env "$@" ncu --set full --clock-control none --import-source on -k regex:smpc_solve_kernel --launch-skip 3 --launch-count 1 \
  -f -o $out/prof_$tag python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --legs none \
  > $out/ncu_$tag.log 2>&1
rc=$?
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/raw_$tag.csv 2>/dev/null
[ "$KEEP_REP" = "1" ] || rm -f $out/prof_$tag.ncu-rep
echo "$tag rc=$rc"
