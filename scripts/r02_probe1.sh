#!/bin/bash
# round 2, first look at the warp-per-problem solver: variants + slice lengths + one ncu capture
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh default "default|DIB_SLICE_ROUNDS=20" "default|DIB_SLICE_ROUNDS=10" "default|DIB_SLICE_ROUNDS=2" w16 w10x2 w5x4 w2x10 gps1 default
cp gpurun_out/sweep.log gpurun_out/r02_probe1_sweep.jsonl
ncu --set full --clock-control none --import-source on -k regex:frustum_solve -s 1 -c 1 -o gpurun_out/prof_r02a_solve \
    python bench.py --samples-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_solve_r02a.log 2>&1
ls -la gpurun_out/*.ncu-rep
