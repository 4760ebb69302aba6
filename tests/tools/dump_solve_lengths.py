"""Dump what the scheduler could know before a solve (its init, the sample's initial guess, the cost / gradient at the
init) next to how long the solve turned out to be (evaluations), for the bench workload: the data behind the solver's
longest-first scheduling order.  GPU tool; writes an .npz."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from deepi2p_b200 import frustum  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--inits", type=int, default=60)
    ap.add_argument("--first-id", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--is-3d", action="store_true")
    ap.add_argument("--out", default="gpurun_out/solve_lengths.npz")
    a = ap.parse_args()
    n_points = 20480
    xyz, pred = bench.make_host_batch_threads(a.first_id, a.samples, n_points)
    from deepi2p_b200 import synthetic as syn
    smp = syn.make_sample(a.first_id, n_points)
    K, H, W = smp["K"], smp["H"], smp["W"]
    xyz_d = torch.from_numpy(xyz).cuda(); pred_d = torch.from_numpy(pred).cuda()
    out = frustum.register_batch(xyz_d, pred_d, n_points, K, H, W, n_inits=a.inits, seed=a.seed, is_2d=not a.is_3d, return_all=True)
    torch.cuda.synchronize()
    prep = frustum.prepare_batch(xyz_d, pred_d, n_points, a.inits, seed=a.seed)
    init = out["init"]                                     # [S,I,4]
    cost0 = torch.zeros(a.samples, a.inits, dtype=torch.float64)
    gnorm0 = torch.zeros(a.samples, a.inits, dtype=torch.float64)
    for i in range(a.inits):
        x = torch.zeros(a.samples, 6, dtype=torch.float64, device="cuda")
        if a.is_3d:
            x[:, 1] = init[:, i, 0]; x[:, 3:6] = init[:, i, 1:4]
        else:
            x[:, 0] = init[:, i, 0]; x[:, 1:4] = init[:, i, 1:4]
        c, g, _ = frustum.evaluate_batch(prep["xyz"], prep["label"], prep["n_pts"], K, x, H, W, not a.is_3d, slice_rounds=0)
        cost0[:, i] = c.cpu(); gnorm0[:, i] = g.abs().amax(dim=1).cpu()
    np.savez_compressed(a.out, init=init.cpu().numpy(), stats=out["stats"].cpu().numpy(), costs=out["costs"].cpu().numpy(),
                        init_y_angle=out["init_y_angle"].cpu().numpy(), n_pts=out["n_pts"].cpu().numpy(),
                        cost0=cost0.numpy(), gnorm0=gnorm0.numpy())
    st = out["stats"].cpu().numpy()
    print("saved", a.out, "evals mean %.1f max %d" % (st[..., 1].mean(), st[..., 1].max()))


if __name__ == "__main__":
    main()
