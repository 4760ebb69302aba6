"""Find (sample, init) solves where the GPU and the CPU oracle disagree, and localise the cause."""
import sys, os, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from concurrent.futures import ThreadPoolExecutor
from deepi2p_b200 import frustum, synthetic as syn

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 5
I = 60
bad = []
for sid in range(ns):
    smp = syn.make_sample(sid)
    iy, pf, lf, _ = oracle.initial_guess(smp["points"], smp["pred"])
    ry, t = syn.make_inits(sid, iy, I)
    xyz, lab, npd = frustum.pack_clouds(pf, lf)
    init = np.concatenate([ry[:, None], t], axis=1)[None]
    g = frustum.solve_batch(xyz, lab, npd, smp["K"], init, smp["H"], smp["W"], return_all=True)
    gp = g["params"][0].cpu().numpy(); gs = g["stats"][0].cpu().numpy(); gc = g["costs"][0].cpu().numpy()
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        outs = list(ex.map(lambda i: oracle.solve(pf, lf, smp["K"], ry[i], t[i], smp["H"], smp["W"], syn.T_LB, syn.T_UB, 500, True, want_residuals=False), range(I)))
    for i in range(I):
        op = outs[i][4]; st = outs[i][3]
        d = np.abs(gp[i, :4] - op[:4]).max()
        if d > 1e-7 or gs[i, 0] != st["iterations"] or gs[i, 1] != st["unique_evals"]:
            bad.append((sid, i, d))
            print("sample %d init %d: max|dx| %.3e  gpu it/ev/ls/term %s  oracle %d/%d/%d/%d  cost gpu %.10f oracle %.10f" % (
                sid, i, d, gs[i].tolist(), st["iterations"], st["unique_evals"], st["line_search_steps"], st["termination"], gc[i], outs[i][1]))
    # evaluation parity at the init and final poses for this sample
    for i in ([b[1] for b in bad if b[0] == sid][:2] or [0]):
        for name, x4 in (("init", np.array([ry[i], *t[i]])), ("gpu-final", gp[i, :4]), ("oracle-final", outs[i][4][:4])):
            x = np.zeros((1, 6)); x[0, :4] = x4
            c, gr, A = frustum.evaluate_batch(xyz, lab, npd, smp["K"], x, smp["H"], smp["W"], True)
            co, go, Ao = oracle.evaluate(pf, lf, smp["K"], x4, smp["H"], smp["W"], True)
            print("   eval %-12s init %2d: cost rel %.2e  grad rel %.2e  JtJ rel %.2e" % (
                name, i, abs(c.item() - co) / max(1, abs(co)), np.abs(gr[0].cpu().numpy() - go).max() / np.abs(go).max(),
                np.abs(A[0].cpu().numpy() - Ao).max() / np.abs(Ao).max()))
print("disagreeing solves:", len(bad), "of", ns * I)
