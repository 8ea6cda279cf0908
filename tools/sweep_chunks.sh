#!/bin/bash
# Sweep the chunk count of the host-buffer pipeline (SMPC_CHUNKS; 0 = heuristic). usage: sweep_chunks.sh workload "<chunks>"
out=gpurun_out/sweep; mkdir -p $out
wl=$1
for c in $2; do
  tag=${wl}_chunks${c}
  SMPC_CHUNKS=$c python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --latency-calls 0 \
    > $out/$tag.json 2> $out/$tag.err
  python - <<P | tee -a $out/summary.txt
import json
try:
    d=json.loads(open("$out/$tag.json").read().strip().splitlines()[-1])
    print("$wl chunks=$c value=%.3fM ms=%.3f e2e=%.3fM e2e_ms=%.3f"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"]))
except Exception as e: print("$wl chunks=$c FAILED",e)
P
done
