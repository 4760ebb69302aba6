#!/bin/bash
# late problems of a batch: sliced from their first pass (count x slice length sweep), same-call A/B
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh "default|DIB_LATE_PROBLEMS=0" "default|DIB_LATE_PROBLEMS=1500" "default|DIB_LATE_PROBLEMS=2960" "default|DIB_LATE_PROBLEMS=6000" "default|DIB_LATE_PROBLEMS=12000" \
  "default|DIB_LATE_PROBLEMS=2960 DIB_LATE_SLICE_ROUNDS=4" "default|DIB_LATE_PROBLEMS=6000 DIB_LATE_SLICE_ROUNDS=4" "default|DIB_LATE_PROBLEMS=12000 DIB_LATE_SLICE_ROUNDS=4" \
  "default|DIB_LATE_PROBLEMS=6000 DIB_LATE_SLICE_ROUNDS=1" "default|DIB_LATE_PROBLEMS=0"
cp gpurun_out/sweep.log gpurun_out/r02_probe15_sweep.jsonl
timeout 240 python -m pytest tests/test_frustum_gpu.py -x -q 2>&1 | tail -2
