#!/bin/bash
# usage: scripts/sweep_env.sh VAR v1 v2 ...   (runs the reduced bench with VAR=v for each v)
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 4 --warmup 2 --no-cpu-baseline --samples-per-gpu ${SWEEP_SAMPLES:-512} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err || { echo "RUN FAILED $var=$v"; tail -3 gpurun_out/sweep_tmp.err; continue; }
  python - "$var=$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep_tmp.json")); r = d["roofline"]
print("%-30s reg/s %8.1f  kernel_ms %8.2f %s frac %.3f" % (sys.argv[1], d["value"], r["kernel_ms"], ["%.1f" % v for v in r.get("kernel_ms_all", [])], r["frac"]), flush=True)
PY
done
