#!/bin/bash
# Experiment runner for `make DEV=1` builds (libsmpc_dev.so): benches "tag workload ENV=..." lines read from stdin.
export SMPC_LIB_PATH=$PWD/nav2_social_mpc_controller_b200/libsmpc_dev.so
out=${OUT:-gpurun_out/exp}; mkdir -p $out
run() { tag=$1; wl=$2; shift 2; env "$@" python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --legs none > $out/$tag.json 2> $out/$tag.err; python - <<PY
import json
try:
    d=json.loads(open("$out/$tag.json").read().strip().split("\n")[-1]); print("$tag", "%.4g"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e: print("$tag ERR", e, open("$out/$tag.err").read()[-300:])
PY
}
while read -r line; do [ -n "$line" ] && run $line; done
