#!/bin/bash
mkdir -p gpurun_out
for lib in "" impre3 impre2 impre1 gk4 gk3 gk2 ""; do
  o=""; [ -n "$lib" ] && o=deepi2p_b200/lib/variants/$lib.so
  DIB_LIB_OVERRIDE=$o timeout 200 python bench.py --ops-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['ops']; print('%-8s index_max %.2f us frac %.3f  ball_query %.2f' % ('$lib' or 'default', d['index_max']['us'], d['index_max']['frac'], d['ball_query']['us']))"
done
for lib in gk4 gk2; do
DIB_LIB_OVERRIDE=deepi2p_b200/lib/variants/$lib.so timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k index_max 2>&1 | tail -2
done
