"""How the exact-path work of a pass is spread over the cloud's rounds (1024 points) and slices: active points
(label 1 outside the image, label 0 inside it) per round of the (label, Morton)-sorted cloud that the solver reads, at
the poses a real solve visits.  GPU tool (uses prepare + a traced solve); the activity test itself is numpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepi2p_b200 import frustum, synthetic as syn  # noqa: E402


def active_mask(xyz, lab, K, H, W, x):
    ry, t = x[0], x[1:4]
    c, s = np.cos(ry), np.sin(ry)
    X = c * xyz[0] + s * xyz[2] + t[0]
    Y = xyz[1] + t[1]
    Z = -s * xyz[0] + c * xyz[2] + t[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = K[0, 0] * X / Z + K[0, 2]
        v = K[1, 1] * Y / Z + K[1, 2]
    inside = (Z > 0) & (u > 0) & (u < W - 1) & (v > 0) & (v < H - 1)
    return ((lab == 1) & ~inside) | ((lab == 0) & inside)


def main():
    smp = syn.make_sample(700)
    xyz_in, pred_in, _ = frustum.pack_clouds(smp["points"][None], smp["pred"][None])
    K, H, W = smp["K"], smp["H"], smp["W"]
    I = 8
    prep = frustum.prepare_batch(xyz_in, pred_in, 20480, I, seed=3)
    out = frustum.solve_batch(prep["xyz"], prep["label"], prep["n_pts"], K, prep["init"], H, W, syn.T_LB, syn.T_UB, 500, True,
                              return_all=True, trace_cap=300)
    torch.cuda.synchronize()
    xyz = prep["xyz"][0].cpu().numpy().astype(np.float64); lab = prep["label"][0].cpu().numpy(); n = int(prep["n_pts"][0])
    xyz, lab = xyz[:, :n], lab[:n]
    rounds = (n + 1023) // 1024
    tr = out["trace"].cpu().numpy()[0]                     # [I,cap,16]
    print("n %d rounds %d; label-1 points %d (rounds 0..%d hold label 0 first? first label-1 index %d)" % (n, rounds, (lab == 1).sum(), rounds - 1, int(np.argmax(lab == 1))))
    for i in range(min(I, 4)):
        valid = tr[i, :, 15] == 1.0
        recs = tr[i][valid]
        for which in (0, len(recs) // 2, len(recs) - 1):
            x = recs[which, :4]
            a = active_mask(xyz, lab, np.asarray(K, dtype=np.float64).reshape(3, 3), H, W, x)
            per_round = np.add.reduceat(a.astype(np.int64), np.arange(0, n, 1024))
            per_slice2 = per_round.reshape(-1, 2).sum(1) if rounds % 2 == 0 else per_round
            print("init %d eval %3d: active %5d (%.1f %%)  per round %s  | 2-round slices: max %d = %.2f x mean; 1-round: max %d = %.2f x mean"
                  % (i, which, a.sum(), 100.0 * a.mean(), per_round.tolist(), per_slice2.max(), per_slice2.max() / max(per_slice2.mean(), 1e-9),
                     per_round.max(), per_round.max() / max(per_round.mean(), 1e-9)))


if __name__ == "__main__":
    main()
