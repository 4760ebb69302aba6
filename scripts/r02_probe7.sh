#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
def show(path):
    try:
        d = json.load(open(path))
        r = d["roofline"]
        print(path, "value %.1f (serial %.1f) e2e %.1f (x%.3f) ms %.2f serial_ms %.2f kernel_ms %.2f frac %.3f tail %.2f" % (d["value"], d["serial"]["value"], d["e2e"]["value"], d["e2e"]["vs_resident"], d["ms_per_step"], d["serial"]["ms_per_step"], r["kernel_ms"], r["frac"], r["tail_ms"]), d["clocks"])
        for k, v in d.get("configs", {}).items():
            if "error" in v: print(" ", k, "ERROR", v["error"]); continue
            if k == "single_sample_60_calls":
                print(" ", k, "dropin ms/registration %.1f (%.2f ms/call); register_batch S1: %.2f ms, %.1f reg/s" % (v["dropin_60_sequential_solvePGivenK"]["ms_per_registration"], v["dropin_60_sequential_solvePGivenK"]["ms_per_call"], v["register_batch_S1_I60"]["ms_per_step"], v["register_batch_S1_I60"]["value"]))
            elif k == "ops_config3":
                print(" ", k, {kk: (round(vv["us"], 1), round(vv.get("frac", vv.get("frac_algorithmic", 0)) or 0, 3)) for kk, vv in v.items() if isinstance(vv, dict) and "us" in vv})
            else:
                print(" ", k, "value %.1f ms %.2f kernel %.2f frac %.3f tail %.2f" % (v["value"], v["ms_per_step"], v["kernel_ms"], v["frac"], v["tail_ms"]))
        print("  cpu", d.get("cpu_baseline", {}).get("value"), "parity", {k: v for k, v in d.get("parity", {}).items() if k not in ("note", "gate", "path")})
    except Exception as e:
        print(path, "parse failed:", e)
show("gpurun_out/r02g_bench.json")
PY
DIB_INDEX_MAX_STREAMING=1 python bench.py --ops-only > gpurun_out/r02g_ops_streaming.json 2>> gpurun_out/r02g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_ops_streaming.json'))['ops']; print('streaming index_max us', d['index_max']['us'])"
ncu --set full --clock-control none --import-source on -k regex:index_max_sorted\|ball_query_split -s 10 -c 3 -o gpurun_out/prof_r02g_ops \
    python bench.py --ops-only > gpurun_out/ncu_ops_r02g.log 2>&1
ls -la gpurun_out/prof_r02g_ops.ncu-rep
