#!/bin/bash
# round 2: scheduling-chunk x slice-length sweep (env knobs only) + the rewritten index_max / ball_query
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -5
python bench.py --ops-only > gpurun_out/r02b_ops.json 2> gpurun_out/r02b_ops.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_ops.json"))["ops"]
for k, v in d.items():
    if isinstance(v, dict) and "us" in v: print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("note",)})
PY
scripts/ab_prebuilt.sh "default|DIB_CHUNK_SAMPLES=512" "default|DIB_CHUNK_SAMPLES=256" "default|DIB_CHUNK_SAMPLES=171" "default|DIB_CHUNK_SAMPLES=128" "default|DIB_CHUNK_SAMPLES=64" "default|DIB_CHUNK_SAMPLES=512 DIB_SLICE_ROUNDS=20" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_ROUNDS=20" "default|DIB_CHUNK_SAMPLES=128 DIB_SLICE_ROUNDS=20" "default|DIB_CHUNK_SAMPLES=512 DIB_SLICE_ROUNDS=10" default
cp gpurun_out/sweep.log gpurun_out/r02_probe2_sweep.jsonl
