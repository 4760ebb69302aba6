#!/bin/bash
# within-call A/B: scripts/ab.sh "<nvcc flags>|<ENV=val ...>" ...
mkdir -p gpurun_out
for item in "$@"; do
  flags="${item%%|*}"; envs="${item#*|}"
  DIB_NVCC_EXTRA="$flags" python -m deepi2p_b200.build --force > /dev/null 2>&1 || { echo "BUILD FAILED: $flags"; continue; }
  env $envs timeout 150 python bench.py ${BENCH_ARGS:-} --steps 4 --warmup 2 --no-cpu-baseline --no-configs --samples-per-gpu ${SWEEP_SAMPLES:-512} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err || { echo "RUN FAILED: $item"; tail -3 gpurun_out/sweep_tmp.err; continue; }
  python - "$item" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep_tmp.json")); r = d["roofline"]
print("%-60s reg/s %8.1f  kernel_ms %8.2f %s frac %.3f" % (sys.argv[1], d["value"], r["kernel_ms"], ["%.1f" % v for v in r.get("kernel_ms_all", [])], r["frac"]), flush=True)
with open("gpurun_out/sweep.log", "a") as f:
    f.write(json.dumps({"cfg": sys.argv[1], "value": d["value"], "kernel_ms": r["kernel_ms"], "kernel_ms_all": r.get("kernel_ms_all"), "frac": r["frac"]}) + "\n")
PY
done
python -m deepi2p_b200.build --force > /dev/null 2>&1
