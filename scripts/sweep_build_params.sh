#!/bin/bash
# Build-time parameter sweep of the solver kernel on the GPU box (run under gpurun).
# usage: scripts/sweep_build_params.sh "<nvcc -D flags>" ["<flags>" ...]
mkdir -p gpurun_out
for cfg in "$@"; do
  DIB_NVCC_EXTRA="$cfg" python -m deepi2p_b200.build --force > /dev/null 2>&1 || { echo "BUILD FAILED: $cfg"; continue; }
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline --samples-per-gpu ${SWEEP_SAMPLES:-256} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err || { echo "RUN FAILED: $cfg"; tail -3 gpurun_out/sweep_tmp.err; continue; }
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep_tmp.json"))
r = d["roofline"]
print("%-70s reg/s %8.1f  kernel_ms %8.2f %s frac %.3f  pt-evals/s %.3e" % (sys.argv[1], d["value"], r["kernel_ms"], ["%.1f" % v for v in r.get("kernel_ms_all", [])], r["frac"], r["point_evals_per_s"]), flush=True)
with open("gpurun_out/sweep.log", "a") as f:
    f.write(json.dumps({"cfg": sys.argv[1], "value": d["value"], "kernel_ms": r["kernel_ms"], "frac": r["frac"]}) + "\n")
PY
done
python -m deepi2p_b200.build --force > /dev/null 2>&1
