#!/bin/bash
# config 2 (4096 clouds x 1 init, 1.4 waves): when should passes start to be sliced?
mkdir -p gpurun_out
for e in "" "DIB_SLICE_AFTER=16" "DIB_SLICE_AFTER=32" "DIB_SLICE_AFTER=48" "DIB_SLICE_AFTER=32 DIB_LATE_PROBLEMS=1136" "DIB_SLICE_AFTER=32 DIB_LATE_PROBLEMS=0" ""; do
  env $e timeout 200 python bench.py --workload single_init --steps 4 --warmup 2 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-44s value %.0f kernel %.2f %s frac %.3f tail %.2f' % ('$e' or 'default', d['value'], r['kernel_ms'], ['%.1f'%v for v in r['kernel_ms_all']], r['frac'], r['tail_ms']))"
done
