#!/bin/bash
# round 2: all GPU tests on the current tree, slice-after x chunk sweep, ops bench + ncu of the ops
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --ops-only > gpurun_out/r02c_ops.json 2> gpurun_out/r02c_ops.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02c_ops.json"))["ops"]
    for k, v in d.items():
        if isinstance(v, dict) and "us" in v: print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("note",)})
except Exception as e:
    print("ops bench failed", e); print(open("gpurun_out/r02c_ops.err").read()[-2000:])
PY
scripts/ab_prebuilt.sh default "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=0" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=32" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=48" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=64" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=96" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=100000" "default|DIB_CHUNK_SAMPLES=512 DIB_SLICE_AFTER=48" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=48 DIB_SLICE_ROUNDS=2" "default|DIB_CHUNK_SAMPLES=256 DIB_SLICE_AFTER=48"
cp gpurun_out/sweep.log gpurun_out/r02_probe3_sweep.jsonl
ncu --set full --clock-control none --import-source on -k regex:index_max_kernel\|ball_query -s 12 -c 2 -o gpurun_out/prof_r02c_ops \
    python bench.py --ops-only > gpurun_out/ncu_ops_r02c.log 2>&1
ls -la gpurun_out/*.ncu-rep
