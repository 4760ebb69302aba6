#!/bin/bash
# Run under gpurun on ONE GPU: tests, smoke, bench lines, ncu launch list and full captures.
# Outputs go to gpurun_out/ (summarise into profiles/ afterwards with scripts/ncu_summary.py).
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --ops > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload single_init --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_single_init.json 2>> gpurun_out/bench_${tag}.err
python bench.py --is-3d --steps 3 --warmup 3 --no-cpu-baseline --samples-per-gpu 256 > gpurun_out/bench_${tag}_6dof.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:frustum_solve -s 1 -c 1 -o gpurun_out/prof_${tag}_solve \
    python bench.py --samples-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_solve_${tag}.log 2>&1
[ -n "$SKIP_OPS_NCU" ] || ncu --set full --clock-control none --import-source on -k regex:index_max_kernel\|ball_query_kernel -s 8 -c 2 -o gpurun_out/prof_${tag}_ops \
    python bench.py --ops-only > gpurun_out/ncu_ops_${tag}.log 2>&1
python - <<PY
import json
for f in ("bench_${tag}.json", "bench_${tag}_reference.json", "bench_${tag}_single_init.json", "bench_${tag}_6dof.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
        r = d.get("roofline", {})
        print(f, "value %.1f e2e %.1f frac %s kernel_ms %s" % (d["value"], d["e2e"]["value"], r.get("frac"), r.get("kernel_ms")))
        if "ops" in d: print("  ops:", {k: (round(v["us"], 1), round(v.get("frac", v.get("frac_algorithmic")) or 0, 3)) for k, v in d["ops"].items() if isinstance(v, dict) and "us" in v})
        if "parity" in d: print("  parity:", {k: v for k, v in d["parity"].items() if k not in ("note", "gate")})
        if "cpu_baseline" in d: print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out | tail -20
