#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -5 gpurun_out/r02d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02d_bench.json"))
    r = d["roofline"]
    print("value %.1f e2e %.1f (x%.3f) ms %.2f kernel_ms %.2f frac %.3f tail %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["vs_resident"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["tail_ms"]))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        if k == "single_sample_60_calls":
            print(k, "dropin ms/registration %.1f (%.2f ms/call); register_batch S1: %.2f ms, %.1f reg/s" % (v["dropin_60_sequential_solvePGivenK"]["ms_per_registration"], v["dropin_60_sequential_solvePGivenK"]["ms_per_call"], v["register_batch_S1_I60"]["ms_per_step"], v["register_batch_S1_I60"]["value"]))
        elif k == "ops_config3":
            print(k, {kk: round(vv["us"], 1) for kk, vv in v.items() if isinstance(vv, dict) and "us" in vv})
        else:
            print(k, "value %.1f ms %.2f kernel %.2f frac %.3f tail %.2f" % (v["value"], v["ms_per_step"], v["kernel_ms"], v["frac"], v["tail_ms"]))
    print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("spread"))
    print("parity", {k: v for k, v in d.get("parity", {}).items() if k not in ("note", "gate", "path")})
except Exception as e:
    print("bench parse failed:", e)
PY
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02d_bench_reference.json 2>> gpurun_out/r02d_bench.err; tail -c 600 gpurun_out/r02d_bench_reference.json; tail -4 gpurun_out/r02d_bench.err
( time python tests/tools/trace_divergence.py --samples 24 --inits 60 --out gpurun_out/r02_trace_divergence ) > gpurun_out/r02_trace.log 2>&1; tail -c 2500 gpurun_out/r02_trace.log
