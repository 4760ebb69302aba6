#!/bin/bash
# within-call A/B over PREBUILT library variants (python -m deepi2p_b200.build --out NAME, with DIB_NVCC_EXTRA set):
#   scripts/ab_prebuilt.sh default pipe2 pipe1 ...      ("default" = the in-tree library)
# Building here instead of on the GPU box keeps nvcc time out of the GPU budget.
mkdir -p gpurun_out
for item in "$@"; do
  name="${item%%|*}"; envs=""; [ "$item" != "$name" ] && envs="${item#*|}"     # "variant|ENV=val ENV2=val"
  lib=""; [ "$name" != default ] && lib="deepi2p_b200/lib/variants/$name.so"
  env DIB_LIB_OVERRIDE="$lib" $envs timeout 150 python bench.py ${BENCH_ARGS:-} --steps 4 --warmup 2 --no-cpu-baseline --no-configs --samples-per-gpu ${SWEEP_SAMPLES:-512} > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_tmp.err || { echo "RUN FAILED: $name"; tail -3 gpurun_out/sweep_tmp.err; continue; }
  python - "$item" <<'PY'
import json, sys
d = json.load(open("gpurun_out/sweep_tmp.json")); r = d["roofline"]
print("%-44s reg/s %8.1f  kernel_ms %8.2f %s frac %.3f tail_ms %s" % (sys.argv[1], d["value"], r["kernel_ms"], ["%.1f" % v for v in r.get("kernel_ms_all", [])], r["frac"], r.get("tail_ms")), flush=True)
with open("gpurun_out/sweep.log", "a") as f:
    f.write(json.dumps({"cfg": "prebuilt:" + sys.argv[1], "value": d["value"], "kernel_ms": r["kernel_ms"], "kernel_ms_all": r.get("kernel_ms_all"), "frac": r["frac"], "tail_ms": r.get("tail_ms")}) + "\n")
PY
done
