#!/bin/bash
# Sweep one environment override of libsmpc over workloads. usage: sweep_env.sh VAR "<values>" workload...
out=gpurun_out/sweep; mkdir -p $out
var=$1; vals=$2; shift 2
for wl in "$@"; do for v in $vals; do
  tag=${wl}_${var}${v}
  env $var=$v python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --latency-calls 0 \
    > $out/$tag.json 2> $out/$tag.err
  python - <<P | tee -a $out/summary.txt
import json
try:
    d=json.loads(open("$out/$tag.json").read().strip().splitlines()[-1])
    print("$wl $var=$v value=%.3fM ms=%.3f e2e=%.3fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6))
except Exception as e: print("$wl $var=$v FAILED",e)
P
done; done
