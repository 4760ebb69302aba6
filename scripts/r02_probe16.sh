#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_frustum_gpu.py -x -q -k "late or reproducible" 2>&1 | tail -3
for i in 1 2 3; do
timeout 200 python bench.py --ops-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['ops']; print('ball_query %.2f us frac %.3f  index_max %.2f  xyz %.2f' % (d['ball_query']['us'], d['ball_query']['frac_algorithmic'], d['index_max']['us'], d['ball_query_xyz']['us']))"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ball_query_split -s 8 -c 1 -o gpurun_out/prof_r02b_ball_query python bench.py --ops-only > gpurun_out/ncu_ops_r02b.log 2>&1
ls -la gpurun_out/prof_r02b_ball_query.ncu-rep
