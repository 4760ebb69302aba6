"""Where a single drop-in solvePGivenK call (numpy in, numpy out) spends its time.  GPU tool."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402  (initial guess / inits as the reference driver makes them)
from deepi2p_b200 import frustum, synthetic as syn  # noqa: E402


def main():
    smp = syn.make_sample(3)
    iy, pts, lab, _ = oracle.initial_guess(smp["points"], smp["pred"])
    ry, t = syn.make_inits(3, iy, 60)
    K, H, W = smp["K"], smp["H"], smp["W"]
    pts = np.asarray(pts, dtype=np.float64); lab = np.asarray(lab)
    def sync(): torch.cuda.synchronize()
    for rep in range(3):
        acc = dict(pack=0.0, solve=0.0, resid=0.0, d2h=0.0, total=0.0)
        for i in range(60):
            sync(); t0 = time.perf_counter()
            xyz, l8, n_pts = frustum.pack_clouds(pts, lab); sync(); t1 = time.perf_counter()
            init = torch.tensor([[[float(ry[i]), t[i][0], t[i][1], t[i][2]]]], dtype=torch.float64)
            out = frustum.solve_batch(xyz, l8, n_pts, np.asarray(K, dtype=np.float64), init, H, W, syn.T_LB, syn.T_UB, 500, True, return_all=True)
            sync(); t2 = time.perf_counter()
            res = frustum.residuals(xyz[0], l8[0], pts.shape[1], np.asarray(K, dtype=np.float64), out["params"][0, 0], H, W, True, host_labels=lab)
            sync(); t3 = time.perf_counter()
            P = out["P"][0].cpu().numpy(); c = float(out["cost"][0].item()); r = res.cpu().numpy()
            t4 = time.perf_counter()
            acc["pack"] += t1 - t0; acc["solve"] += t2 - t1; acc["resid"] += t3 - t2; acc["d2h"] += t4 - t3; acc["total"] += t4 - t0
        ev = None
        print("rep %d per call: " % rep + "  ".join("%s %.3f ms" % (k, v / 60 * 1e3) for k, v in acc.items()))
    # the solve alone, sorted vs unsorted cloud, kernel time by CUDA events
    xyz, l8, n_pts = frustum.pack_clouds(pts, lab)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    init = torch.tensor([[[float(ry[0]), t[0][0], t[0][1], t[0][2]]]], dtype=torch.float64)
    sx, sl, sn = frustum.sort_clouds(xyz, l8, pts.shape[1])
    for _ in range(3):
        e0.record(); frustum.sort_clouds(xyz, l8, pts.shape[1]); e1.record(); sync()
    print("sort_clouds device time %.3f ms" % e0.elapsed_time(e1))
    for name, (a, b, c) in (("unsorted", (xyz, l8, n_pts)), ("sorted", (sx, sl, sn))):
        for _ in range(3):
            e0.record(); out = frustum.solve_batch(a, b, c, np.asarray(K, dtype=np.float64), init, H, W, syn.T_LB, syn.T_UB, 500, True, return_all=True); e1.record(); sync()
        print(name, "solve_batch device time %.3f ms, evaluations %d" % (e0.elapsed_time(e1), int(out["stats"][0, 0, 1])))


if __name__ == "__main__":
    main()
