"""Per-pass latency of a helped problem, from a library built with -DDIB_PASS_TIMING (trace slots 13 / 14 then hold the
nanoseconds from open_pass to "all slices in" and the nanoseconds of the control step that follows).  GPU tool:
DIB_LIB_OVERRIDE=deepi2p_b200/lib/variants/timing.so python tests/tools/pass_timing.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepi2p_b200 import frustum, synthetic as syn  # noqa: E402


def main():
    for S, I in ((1, 1), (1, 60), (8, 60)):
        smps = [syn.make_sample(700 + s) for s in range(S)]
        xyz_in, pred_in, _ = frustum.pack_clouds(np.stack([s["points"] for s in smps]), np.stack([s["pred"] for s in smps]))
        K, H, W = smps[0]["K"], smps[0]["H"], smps[0]["W"]
        prep = frustum.prepare_batch(xyz_in, pred_in, 20480, I, seed=3)
        out = None
        for _ in range(3):
            out = frustum.solve_batch(prep["xyz"], prep["label"], prep["n_pts"], K, prep["init"], H, W, syn.T_LB, syn.T_UB, 500, True,
                                      return_all=True, trace_cap=300)
        torch.cuda.synchronize()
        tr = out["trace"].cpu().numpy()                     # [S,I,cap,16]
        valid = tr[..., 15] == 1.0
        pass_ns = tr[..., 13][valid]; ctl_ns = tr[..., 14][valid]
        ev = out["stats"].cpu().numpy()[..., 1]
        print("S %d I %d: passes %d (max per solve %d)  open->complete: mean %.1f us p50 %.1f p90 %.1f | control step: mean %.1f us p50 %.1f p90 %.1f"
              % (S, I, valid.sum(), ev.max(), pass_ns.mean() / 1e3, np.percentile(pass_ns, 50) / 1e3, np.percentile(pass_ns, 90) / 1e3,
                 ctl_ns.mean() / 1e3, np.percentile(ctl_ns, 50) / 1e3, np.percentile(ctl_ns, 90) / 1e3))
        # split the control step by what it did: a line-search interpolation shows as phase 1 records
        ph = tr[..., 10][valid]
        for p in (0, 1, 2):
            m = ph == p
            if m.any():
                print("    phase %d: n %d  pass %.1f us  control %.1f us" % (p, m.sum(), pass_ns[m].mean() / 1e3, ctl_ns[m].mean() / 1e3))


if __name__ == "__main__":
    main()
