#!/bin/bash
# Sweep the launch-mapping overrides for workloads: lanes per problem (SMPC_GROUP), warps per CTA (SMPC_WARPS) and
# resident 4-warp CTAs per SM (SMPC_MINB). usage: sweep_groups.sh "<groups>" "<warps>" "<minbs>" workload...
out=gpurun_out/sweep; mkdir -p $out
groups=$1; warps=$2; minbs=$3; shift 3
for wl in "$@"; do
  for g in $groups; do for w in $warps; do for mb in $minbs; do
    tag=${wl}_g${g}_w${w}_mb${mb}
    SMPC_GROUP=$g SMPC_WARPS=$w SMPC_MINB=$mb python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline \
      --latency-calls 0 > $out/$tag.json 2> $out/$tag.err
    python - <<P | tee -a $out/summary.txt
import json
try:
    d=json.loads(open("$out/$tag.json").read().strip().splitlines()[-1])
    print("$wl G=$g W=$w MB=$mb value=%.3fM ms=%.3f e2e=%.3fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6))
except Exception as e: print("$wl G=$g W=$w MB=$mb FAILED",e)
P
  done; done; done
done
