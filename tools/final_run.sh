#!/bin/bash
# Round evidence run on ONE GPU of the box: the driver's default bench command and its reference arm, a bench line per
# other workload, single-solve latency, the pre-solve / fleet-tick measurements, the per-problem flip logs, the ncu
# launch list of the default bench command and one `ncu --set full` capture per profiled workload (raw page exported to
# CSV on the box; only two .ncu-rep files travel back — gpurun_out/ is capped at 64 MiB per call).
# usage: tools/final_run.sh OUTDIR        afterwards, here: python tools/summarize_profiles.py OUTDIR r02
out=${1:-gpurun_out/final}; mkdir -p $out
python bench.py > $out/bench_default.json 2>> $out/err.log || echo "default bench failed" >> $out/err.log
cp $out/bench_default.json $out/bench_soc_work_obst_x65536_A20.json
python bench.py --impl reference --steps 3 --warmup 1 > $out/ref_default.json 2>> $out/err.log
for wl in obst_only_x4096 obst_only_x65536 soc_work_obst_x16384_A3 soc_work_obst_x65536_A3 multistart_256x1024 \
          crowd_x16384_A50; do
  python bench.py --workload $wl --legs none > $out/bench_$wl.json 2>> $out/err.log || echo "bench $wl failed" >> $out/err.log
done
python tools/latency.py > $out/latency.json 2>> $out/err.log
python tools/bench_presolve.py > $out/presolve.json 2>> $out/err.log
for spec in corridor:3072 crowd_A3:1536 crowd_A20:512 crowd_A50:256 blocks18:64; do
  wl=${spec%%:*}; n=${spec##*:}
  python tools/flip_log.py --workload $wl --n $n --out $out/flip_log_$wl.json > $out/flip_$wl.log 2>&1
  [ $wl = corridor ] || [ $wl = blocks18 ] || \
    python tools/flip_log.py --workload $wl --n $n --ceres-compat 220 --out $out/flip_log_${wl}_ceres220.json > $out/flip220_$wl.log 2>&1
done
# launch list of the default bench command (every kernel launch once, serialised, cold cache: shares, not absolutes)
# (13 minutes of box time: the two 10^6-problem legs run their kernels serialised; SKIP_LAUNCH_LIST=1 leaves it out)
[ "$SKIP_LAUNCH_LIST" = "1" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
KEEP_REP=1 tools/ncu_capture.sh $out soc_work_obst_x65536_A20 soc_work_obst_x65536_A20
KEEP_REP=1 tools/ncu_capture.sh $out obst_only_x4096 obst_only_x4096
for wl in obst_only_x65536 soc_work_obst_x16384_A3 crowd_x16384_A50; do tools/ncu_capture.sh $out $wl $wl; done
tail -n 5 $out/err.log
