#!/bin/bash
# team shape 20 warps x 1 CTA with 20 one-round slices for small batches / late problems, against the shipped 10 x 2
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
scripts/ab_prebuilt.sh default w20x1 "w20x1s20|DIB_SMALL_SLICE_ROUNDS=1" "w20x1s20|DIB_SMALL_SLICE_ROUNDS=2" default "w20x1s20|DIB_SMALL_SLICE_ROUNDS=1"
cp gpurun_out/sweep.log gpurun_out/r02_probe18_sweep.jsonl
for v in "default:2" "w20x1s20:1" "w20x1s20:2"; do
  name=${v%%:*}; r=${v#*:}; lib=""; [ "$name" != default ] && lib=deepi2p_b200/lib/variants/$name.so
  DIB_LIB_OVERRIDE=$lib DIB_SMALL_SLICE_ROUNDS=$r timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --samples-per-gpu 512 > gpurun_out/p18_$name$r.json 2> gpurun_out/p18_$name$r.err || { echo "configs run failed $v"; tail -3 gpurun_out/p18_$name$r.err; continue; }
  python - $name$r <<'PY'
import json, sys
d = json.load(open("gpurun_out/p18_%s.json" % sys.argv[1])); c = d["configs"]
print("%s value %.1f serial %.1f kernel %.2f tail %.2f" % (sys.argv[1], d["value"], d["serial"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["tail_ms"]))
print("   cfg1 call %.2f ms  S1 %.2f ms | cfg2 %.0f/s kernel %.2f frac %.3f | 6dof %.1f kernel %.2f frac %.3f" % (
    c["single_sample_60_calls"]["dropin_60_sequential_solvePGivenK"]["ms_per_call"], c["single_sample_60_calls"]["register_batch_S1_I60"]["ms_per_step"],
    c["single_init_4096"]["value"], c["single_init_4096"]["kernel_ms"], c["single_init_4096"]["frac"],
    c["sixdof"]["value"], c["sixdof"]["kernel_ms"], c["sixdof"]["frac"]))
PY
done
DIB_LIB_OVERRIDE=deepi2p_b200/lib/variants/w20x1s20.so DIB_SMALL_SLICE_ROUNDS=1 timeout 300 python -m pytest tests/test_frustum_gpu.py -x -q 2>&1 | tail -2
