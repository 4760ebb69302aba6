#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh default w10x2 w5x4 w4x5 default w10x2
cp gpurun_out/sweep.log gpurun_out/r02_probe12_sweep.jsonl
