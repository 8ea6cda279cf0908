#!/bin/bash
# Sweep the launch-mapping overrides for workloads: lanes per problem (SMPC_GROUP) and warps per CTA (SMPC_WARPS:
# 4 / 16 with people, 4 / 12 without). usage: sweep_groups.sh "<groups>" "<warps>" workload...
out=gpurun_out/sweep; mkdir -p $out
groups=$1; warps=$2; shift 2
for wl in "$@"; do
  for g in $groups; do for w in $warps; do
    tag=${wl}_g${g}_w${w}
    SMPC_GROUP=$g SMPC_WARPS=$w python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline \
      --latency-calls 0 > $out/$tag.json 2> $out/$tag.err
    python - <<P | tee -a $out/summary.txt
import json
try:
    d=json.loads(open("$out/$tag.json").read().strip().splitlines()[-1])
    print("$wl G=$g W=$w value=%.3fM ms=%.3f e2e=%.3fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6))
except Exception as e: print("$wl G=$g W=$w FAILED",e)
P
  done; done
done
