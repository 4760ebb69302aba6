#!/bin/bash
# Round-2 final evidence, one GPU: tests, smoke, both bench arms, ncu launch list + full captures, parity study.
# Outputs go to gpurun_out/; scripts/r02_summarise.sh (run in the build container) turns them into profiles/.
tag=${1:-r02final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_${tag}.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${tag}.csv \
    timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 900 ncu --set full --metrics lts__t_bytes.sum --clock-control none --import-source on -k regex:frustum_solve -s 3 -c 1 -o gpurun_out/prof_${tag}_solve \
    python bench.py --samples-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_solve_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_max_kernel -s 8 -c 1 -o gpurun_out/prof_${tag}_index_max \
    python bench.py --ops-only > gpurun_out/ncu_ops_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ball_query_split -s 8 -c 1 -o gpurun_out/prof_${tag}_ball_query \
    python bench.py --ops-only >> gpurun_out/ncu_ops_${tag}.log 2>&1
timeout 900 python tests/tools/trace_divergence.py --samples 24 --inits 60 --out gpurun_out/${tag}_trace_divergence > gpurun_out/trace_${tag}.log 2>&1
timeout 900 python tests/tools/trace_divergence.py --samples 8 --inits 60 --is-3d --hybrid-in-gate 40 --out gpurun_out/${tag}_trace_divergence_6dof > gpurun_out/trace_${tag}_6dof.log 2>&1
python - <<PY
import json
for f in ("bench_${tag}.json", "bench_${tag}_reference.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
        r = d.get("roofline", {})
        print(f, "value %.1f e2e %.1f frac %s kernel_ms %s" % (d["value"], d["e2e"]["value"], r.get("frac"), r.get("kernel_ms")))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out | tail -20
