#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_frustum_gpu.py tests/test_fork_dropin_gpu.py -x -q -k "sort_clouds or dropin or fork" 2>&1 | tail -3
timeout 200 python tests/tools/dropin_breakdown.py 2>&1 | tail -8
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --samples-per-gpu 512 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['single_sample_60_calls']
print('value %.1f kernel %.2f | dropin ms_per_call %.3f ms_per_registration %.1f | S1 %.2f ms' % (d['value'], d['roofline']['kernel_ms'], c['dropin_60_sequential_solvePGivenK']['ms_per_call'], c['dropin_60_sequential_solvePGivenK']['ms_per_registration'], c['register_batch_S1_I60']['ms_per_step']))"
