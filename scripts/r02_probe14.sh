#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tools/dump_solve_lengths.py --out gpurun_out/solve_lengths.npz 2>&1 | tail -3
for lib in "" deepi2p_b200/lib/variants/ballq_persist.so ""  deepi2p_b200/lib/variants/ballq_persist.so; do
  DIB_LIB_OVERRIDE=$lib timeout 200 python bench.py --ops-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['ops']; print('$lib', 'ball_query %.2f us frac %.3f  index_max %.2f' % (d['ball_query']['us'], d['ball_query']['frac_algorithmic'], d['index_max']['us']))"
done
DIB_LIB_OVERRIDE=deepi2p_b200/lib/variants/ballq_persist.so timeout 300 python -m pytest tests/test_ops_gpu.py -x -q 2>&1 | tail -2
