#!/bin/bash
# with the interleaved layout: re-check the slicing parameters of the large batch
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh default "default|DIB_SLICE_AFTER=32" "default|DIB_SLICE_AFTER=64" "default|DIB_SLICE_AFTER=96" "default|DIB_SLICE_ROUNDS=5" "default|DIB_SLICE_ROUNDS=2" "default|DIB_SLICE_ROUNDS=10" "default|DIB_SLICE_AFTER=64 DIB_SLICE_ROUNDS=2" "default|DIB_LATE_PROBLEMS=1500" default
cp gpurun_out/sweep.log gpurun_out/r02_probe20_sweep.jsonl
