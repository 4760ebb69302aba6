#!/usr/bin/env python
"""GPU vs CPU-oracle parity distribution over many full-size solves (run on the GPU box):
    python tests/tools/parity_distribution.py [n_samples] [6dof] > gpurun_out/parity_distribution.json"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from concurrent.futures import ThreadPoolExecutor
from deepi2p_b200 import frustum, synthetic as syn

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20
is_2d = not (len(sys.argv) > 2 and sys.argv[2] == "6dof")
I, P = 60, (4 if is_2d else 6)
drot, dtr, same, reg_ok, cost_le, cost_rel = [], [], 0, 0, 0, []
for sid in range(1000, 1000 + ns):
    smp = syn.make_sample(sid)
    iy, pf, lf, _ = oracle.initial_guess(smp["points"], smp["pred"])
    ry, t = syn.make_inits(sid, iy, I)
    xyz, lab, npd = frustum.pack_clouds(pf, lf)
    init = np.concatenate([ry[:, None], t], axis=1)[None]
    g = frustum.solve_batch(xyz, lab, npd, smp["K"], init, smp["H"], smp["W"], is_2d=is_2d, return_all=True)
    gp = g["params"][0].cpu().numpy(); gs = g["stats"][0].cpu().numpy(); gc = g["costs"][0].cpu().numpy()
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        outs = list(ex.map(lambda i: oracle.solve(pf, lf, smp["K"], ry[i], t[i], smp["H"], smp["W"], syn.T_LB, syn.T_UB,
                                                  500, is_2d, want_residuals=False), range(I)))
    op = np.stack([o[4] for o in outs]); oc = np.array([o[1] for o in outs])
    nr = P - 3
    drot += list(np.linalg.norm(gp[:, :nr] - op[:, :nr], axis=1)); dtr += list(np.linalg.norm(gp[:, nr:P] - op[:, nr:P], axis=1))
    same += sum(int(gs[i, 0] == outs[i][3]["iterations"] and gs[i, 1] == outs[i][3]["unique_evals"]) for i in range(I))
    bg, bo = int(np.argmin(gc)), int(np.argmin(oc))
    reg_ok += int(np.linalg.norm(gp[bg, :nr] - op[bo, :nr]) < 1e-4 and np.linalg.norm(gp[bg, nr:P] - op[bo, nr:P]) < 1e-3)
    cost_le += int(gc[bg] <= oc[bo] * (1 + 1e-9)); cost_rel.append(float(gc[bg] / oc[bo] - 1))
drot, dtr = np.array(drot), np.array(dtr)
within = (drot < 1e-4) & (dtr < 1e-3)
print(json.dumps({
    "what": "GPU solve_batch vs CPU oracle, %d KITTI-shaped samples x %d inits, 20480 points, %s" % (ns, I, "4-DoF" if is_2d else "6-DoF"),
    "solves": int(within.size), "solves_within_gate": int(within.sum()), "fraction_within_gate": float(within.mean()),
    "solves_identical_iteration_and_evaluation_counts": int(same),
    "rot_rad": {"median": float(np.median(drot)), "p90": float(np.percentile(drot, 90)), "p99": float(np.percentile(drot, 99)), "max": float(drot.max())},
    "trans_m": {"median": float(np.median(dtr)), "p90": float(np.percentile(dtr, 90)), "p99": float(np.percentile(dtr, 99)), "max": float(dtr.max())},
    "registrations": ns, "best_of_60_pose_within_gate": reg_ok, "gpu_best_cost_le_oracle_best_cost": cost_le,
    "best_cost_relative_difference": {"median": float(np.median(cost_rel)), "min": float(np.min(cost_rel)), "max": float(np.max(cost_rel))},
    "gate": "1e-4 rad / 1e-3 m"}))
