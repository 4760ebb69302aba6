#!/bin/bash
# ncu capture of the 6-DoF solver kernel (same workload as the sixdof sub-record of the bench line)
mkdir -p gpurun_out
timeout 900 ncu --set full --metrics lts__t_bytes.sum --clock-control none --import-source on -k regex:frustum_solve -s 3 -c 1 -o gpurun_out/prof_r02_solve6 \
    python bench.py --is-3d --samples-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_solve6.log 2>&1
ls -la gpurun_out/prof_r02_solve6.ncu-rep
