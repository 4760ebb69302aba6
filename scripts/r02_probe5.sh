#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
def show(path):
    try:
        d = json.load(open(path))
        r = d["roofline"]
        print(path, "value %.1f e2e %.1f (x%.3f) ms %.2f kernel_ms %.2f frac %.3f tail %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["vs_resident"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["tail_ms"]))
        for k, v in d.get("configs", {}).items():
            if "error" in v: print(" ", k, "ERROR", v["error"]); continue
            if k == "single_sample_60_calls":
                print(" ", k, "dropin ms/registration %.1f (%.2f ms/call); register_batch S1: %.2f ms, %.1f reg/s" % (v["dropin_60_sequential_solvePGivenK"]["ms_per_registration"], v["dropin_60_sequential_solvePGivenK"]["ms_per_call"], v["register_batch_S1_I60"]["ms_per_step"], v["register_batch_S1_I60"]["value"]))
            elif k == "ops_config3":
                print(" ", k, {kk: round(vv["us"], 1) for kk, vv in v.items() if isinstance(vv, dict) and "us" in vv})
            else:
                print(" ", k, "value %.1f ms %.2f kernel %.2f frac %.3f tail %.2f" % (v["value"], v["ms_per_step"], v["kernel_ms"], v["frac"], v["tail_ms"]))
    except Exception as e:
        print(path, "parse failed:", e)
show("gpurun_out/r02e_bench.json")
PY
DIB_LIB_OVERRIDE=deepi2p_b200/lib/variants/f6w16gps1.so python bench.py --is-3d --samples-per-gpu 256 --steps 3 --warmup 2 --no-cpu-baseline --no-configs > gpurun_out/r02e_6dof_w16.json 2>> gpurun_out/r02e_bench.err
python bench.py --is-3d --samples-per-gpu 256 --steps 3 --warmup 2 --no-cpu-baseline --no-configs > gpurun_out/r02e_6dof_w14.json 2>> gpurun_out/r02e_bench.err
python - <<'PY'
import json
for p in ("gpurun_out/r02e_6dof_w16.json", "gpurun_out/r02e_6dof_w14.json"):
    try:
        d = json.load(open(p)); r = d["roofline"]
        print(p, "value %.1f kernel_ms %.2f frac %.3f tail %.2f" % (d["value"], r["kernel_ms"], r["frac"], r["tail_ms"]))
    except Exception as e: print(p, e)
PY
( time python tests/tools/trace_divergence.py --samples 24 --inits 60 --out gpurun_out/r02_trace_divergence ) > gpurun_out/r02_trace.log 2>&1; tail -c 1500 gpurun_out/r02_trace.log
