"""Writes tests/golden/frustum_small.npz: seeded inputs of the registration path together with the outputs of the
CPU oracle (oracle/frustum_oracle.cpp, the restatement of evaluation/frustum_reg/src/registration.cpp:9-186) and
of oracle.cluster_assign (models/networks_pc.py:60-85).

The reference's own solver cannot produce these vectors here (Ceres / Eigen are not installable offline, SURVEY.md
8c), so this fixture pins the ORACLE -- it guards the checker against regressions and gives the GPU tests a
committed set of vectors that does not depend on the oracle being rebuilt identically on the GPU box.

    python oracle/build.py && python tests/golden/make_frustum_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from deepi2p_b200 import synthetic as syn  # noqa: E402

S, N, I = 3, 1024, 4
out = {"S": np.int32(S), "N": np.int32(N), "I": np.int32(I), "H": np.float64(syn.KITTI["H"]), "W": np.float64(syn.KITTI["W"])}
pts, pred, Ks, x4, x6, inits = [], [], [], [], [], []
ev4, ev6, sol4, sol6 = [], [], [], []
for s in range(S):
    smp = syn.make_sample(4200 + s, n_points=N)
    iy, _, _, _ = oracle.initial_guess(smp["points"], smp["pred"])
    ry, t = syn.make_inits(4200 + s, iy, I)
    a4 = np.array([smp["ry_gt"] + 0.04, smp["t_gt"][0] + 0.25, 0.03, smp["t_gt"][2] - 0.4])
    a6 = np.array([0.015, smp["ry_gt"] + 0.04, -0.02, smp["t_gt"][0] + 0.25, 0.03, smp["t_gt"][2] - 0.4])
    pts.append(smp["points"].astype(np.float32)); pred.append(smp["pred"].astype(np.int8)); Ks.append(smp["K"])
    x4.append(a4); x6.append(a6); inits.append(np.concatenate([ry[:, None], t], axis=1))
    for is_2d, x, ev, sol in ((True, a4, ev4, sol4), (False, a6, ev6, sol6)):
        c, g, A = oracle.evaluate(smp["points"], smp["pred"], smp["K"], x, smp["H"], smp["W"], is_2d)
        ev.append(np.concatenate([[c], g, A.reshape(-1)]))
        ms = oracle.solve_multistart(smp["points"], smp["pred"], smp["K"], ry, t, smp["H"], smp["W"], syn.T_LB, syn.T_UB,
                                     500, is_2d)
        sol.append(np.concatenate([ms["params"], ms["costs"][:, None],
                                   np.array([[st["iterations"], st["unique_evals"], st["termination"]] for st in ms["stats"]],
                                            dtype=np.float64)], axis=1))
out.update(points=np.stack(pts), pred=np.stack(pred), K=np.stack(Ks), x4=np.stack(x4), x6=np.stack(x6),
           inits=np.stack(inits), eval4=np.stack(ev4), eval6=np.stack(ev6), solve4=np.stack(sol4), solve6=np.stack(sol6))
# clustering front-end
rng = np.random.default_rng(77)
cpc = rng.uniform(-30, 30, (2, 3, 600)).astype(np.float32)
cnode = cpc[:, :, rng.permutation(600)[:24]].copy()
cnode[:, :, 5] = cnode[:, :, 2]                    # a duplicated node: exact ties
ca = oracle.cluster_assign(cpc, cnode, 3)
out.update(ca_pc=cpc, ca_node=cnode, ca_min_k_idx=ca["min_k_idx"], ca_count=ca["count"], ca_mean=ca["cluster_mean"],
           ca_decentered=ca["pc_decentered"])
np.savez_compressed(os.path.join(HERE, "frustum_small.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
