#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh default gpipe gps3 gpipe_gps1 default gpipe
cp gpurun_out/sweep.log gpurun_out/r02_probe8_sweep.jsonl
python -m pytest tests/test_fork_dropin_gpu.py -q -s 2>&1 | grep -E "per-call|passed|failed"
compute-sanitizer --tool memcheck python scripts/sanitize_run.py all > gpurun_out/r02_sanitize_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitize_memcheck.log
compute-sanitizer --tool racecheck python scripts/sanitize_run.py solver > gpurun_out/r02_sanitize_racecheck.log 2>&1; tail -6 gpurun_out/r02_sanitize_racecheck.log
