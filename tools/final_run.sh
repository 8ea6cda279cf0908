#!/bin/bash
# Round evidence run on the GPU box: bench lines of every workload (+ the reference arm), single-solve latency, the ncu
# launch list of the default bench command and one `ncu --set full` capture per profiled workload (raw page exported
# to CSV on the box; only the default workload's .ncu-rep travels back — gpurun_out/ is capped at 64 MiB per call).
# usage: tools/final_run.sh OUTDIR
out=${1:-gpurun_out/final}; mkdir -p $out
for wl in obst_only_x4096 obst_only_x65536 soc_work_obst_x16384_A3 soc_work_obst_x65536_A3 soc_work_obst_x65536_A20 \
          multistart_256x1024 crowd_x16384_A50; do
  python bench.py --workload $wl > $out/bench_$wl.json 2>> $out/err.log || echo "bench $wl failed" >> $out/err.log
done
python bench.py --impl reference --steps 3 --warmup 1 > $out/ref_obst_only_x4096.json 2>> $out/err.log
python tools/latency.py > $out/latency.json 2>> $out/err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --legs none > $out/ncu_launches.log 2>&1
KEEP_REP=1 tools/ncu_capture.sh $out obst_only_x4096 obst_only_x4096
for wl in obst_only_x65536 soc_work_obst_x16384_A3 soc_work_obst_x65536_A20; do tools/ncu_capture.sh $out $wl $wl; done
tail -n 3 $out/err.log
