#!/bin/bash
# global board (cross-SM end-of-batch help): correctness first, then same-call A/B against DIB_GLOBAL_HELP=0
# (ran against the experimental kernel of profiles/r02_global_board.patch, applied on commit 926bb8c; the knob is not in the shipped source)
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
timeout 240 python -m pytest tests/test_frustum_gpu.py -x -q 2>&1 | tail -4
scripts/ab_prebuilt.sh "default|DIB_GLOBAL_HELP=0" "default|DIB_GLOBAL_HELP=1" "default|DIB_GLOBAL_HELP=0" "default|DIB_GLOBAL_HELP=1"
cp gpurun_out/sweep.log gpurun_out/r02_probe13_sweep.jsonl
for g in 0 1; do
  DIB_GLOBAL_HELP=$g timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --samples-per-gpu 512 > gpurun_out/p13_g$g.json 2> gpurun_out/p13_g$g.err || { echo "configs run failed g=$g"; tail -3 gpurun_out/p13_g$g.err; }
  python - $g <<'PY'
import json, sys
d = json.load(open("gpurun_out/p13_g%s.json" % sys.argv[1])); c = d["configs"]
print("g=%s value %.1f serial %.1f e2e %.1f kernel %.2f" % (sys.argv[1], d["value"], d["serial"]["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
print("   cfg1 call %.2f ms  S1 %.2f ms | cfg2 %.0f/s kernel %.2f frac %.3f | 6dof %.1f kernel %.2f frac %.3f" % (
    c["single_sample_60_calls"]["dropin_60_sequential_solvePGivenK"]["ms_per_call"], c["single_sample_60_calls"]["register_batch_S1_I60"]["ms_per_step"],
    c["single_init_4096"]["value"], c["single_init_4096"]["kernel_ms"], c["single_init_4096"]["frac"],
    c["sixdof"]["value"], c["sixdof"]["kernel_ms"], c["sixdof"]["frac"]))
PY
done
