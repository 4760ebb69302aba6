#!/usr/bin/env python
"""Extract DRAM / L2 traffic and pipe fractions of the solve kernel from an ncu report into profiles/r02_traffic.json.
   python scripts/ncu_traffic.py gpurun_out/prof_X_solve.ncu-rep <samples_per_gpu> <inits> <is_2d 0|1> [out.json]"""
import csv, io, json, os, subprocess, sys
rep, S, I, is2d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(vals[i].replace(",", "")); u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"kernel": vals[hdr.index("Kernel Name")], "samples_per_gpu": S, "inits": I, "is_2d": bool(is2d),
       "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
       "lts_t_bytes": get("lts__t_bytes.sum") if "lts__t_bytes.sum" in hdr else None,
       "gpu_time_ns": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[units[hdr.index("gpu__time_duration.sum")]],
       "source": os.path.basename(rep)}
def pct(name):
    try:
        return float(vals[hdr.index(name)].replace(",", ""))
    except (ValueError, IndexError):
        return None
# the FP64-ALU fraction is reported next to the bandwidth fraction
out["fp64_pipe_active_pct"] = pct("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
out["issue_active_pct"] = pct("smsp__issue_active.avg.pct_of_peak_sustained_active")
out["warps_active_pct"] = pct("sm__warps_active.avg.pct_of_peak_sustained_active")
out["l2_read_sectors_from_l1"] = pct("lts__t_sectors_srcunit_tex_op_read.sum")
out["dram_throughput_pct"] = pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
out["registers_per_thread"] = pct("launch__registers_per_thread")
out["grid_size"] = pct("launch__grid_size")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = sys.argv[5] if len(sys.argv) > 5 else os.path.join(root, "profiles", "r02_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(out)
