#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
scripts/ab_prebuilt.sh default "default|DIB_PRIO_AFTER=1000000" "default|DIB_PRIO_AFTER=48" "default|DIB_PRIO_AFTER=96" "default|DIB_PRIO_AFTER=32 DIB_SLICE_AFTER=32" "default|DIB_PRIO_AFTER=48 DIB_SLICE_ROUNDS=2" default
cp gpurun_out/sweep.log gpurun_out/r02_probe10_sweep.jsonl
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err
python - <<'PY'
import json
def show(path):
    try:
        d = json.load(open(path))
        r = d["roofline"]
        print(path, "value %.1f (serial %.1f) e2e %.1f (x%.3f) ms %.2f serial_ms %.2f kernel_ms %.2f frac %.3f tail %.2f" % (d["value"], d["serial"]["value"], d["e2e"]["value"], d["e2e"]["vs_resident"], d["ms_per_step"], d["serial"]["ms_per_step"], r["kernel_ms"], r["frac"], r["tail_ms"]), r.get("cta_exit_after_queue_empty_ms"))
        for k, v in d.get("configs", {}).items():
            if "error" in v: print(" ", k, "ERROR", v["error"]); continue
            if k == "single_sample_60_calls":
                print(" ", k, "dropin ms/registration %.1f (%.2f ms/call); register_batch S1: %.2f ms, %.1f reg/s" % (v["dropin_60_sequential_solvePGivenK"]["ms_per_registration"], v["dropin_60_sequential_solvePGivenK"]["ms_per_call"], v["register_batch_S1_I60"]["ms_per_step"], v["register_batch_S1_I60"]["value"]))
            elif k == "ops_config3":
                print(" ", k, {kk: (round(vv["us"], 1)) for kk, vv in v.items() if isinstance(vv, dict) and "us" in vv})
            else:
                print(" ", k, "value %.1f ms %.2f kernel %.2f frac %.3f tail %.2f" % (v["value"], v["ms_per_step"], v["kernel_ms"], v["frac"], v["tail_ms"]))
    except Exception as e:
        print(path, "parse failed:", e)
show("gpurun_out/r02i_bench.json")
PY
