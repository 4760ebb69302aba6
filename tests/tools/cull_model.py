"""CPU model of the solver's box cull: how many 32-point groups are skipped / surely active / undecided at the poses
of a solve, as a function of HOW the cloud is ordered before it is cut into groups (the sort key of
frustum_prepare_kernel).  numpy only; the box test is the kernel's (box_state) in float64 without the fp32 margins."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from deepi2p_b200 import synthetic as syn  # noqa: E402


def spread(v, bits, stride):
    out = np.zeros_like(v)
    for b in range(bits):
        out |= ((v >> b) & 1) << (b * stride)
    return out


def quant(a, bits):
    lo, hi = a.min(), a.max()
    q = ((a - lo) * ((1 << bits) / max(hi - lo, 1e-30))).astype(np.int64)
    return np.clip(q, 0, (1 << bits) - 1)


def keys(p, lab, mode):
    x, y, z = p
    cls = np.where(lab == 0, 0, np.where(lab == 1, 1, 2)).astype(np.int64)
    if mode == "none":
        return np.arange(len(x))
    if mode == "xz6":                     # shipped: 6 + 6 bit Morton of (x, z)
        m = (spread(quant(x, 6), 6, 2) << 1) | spread(quant(z, 6), 6, 2)
    elif mode == "xz7":
        m = (spread(quant(x, 7), 7, 2) << 1) | spread(quant(z, 7), 7, 2)
    elif mode == "xyz4":
        m = (spread(quant(x, 4), 4, 3) << 2) | (spread(quant(y, 4), 4, 3) << 1) | spread(quant(z, 4), 4, 3)
    elif mode == "xyz5":
        m = (spread(quant(x, 5), 5, 3) << 2) | (spread(quant(y, 5), 5, 3) << 1) | spread(quant(z, 5), 5, 3)
    elif mode == "xz5y2":                 # Morton of (x, z) 5 + 5 bits, then 2 bits of height as the LOW bits
        m = (((spread(quant(x, 5), 5, 2) << 1) | spread(quant(z, 5), 5, 2)) << 2) | quant(y, 2)
    elif mode == "xz6y2":
        m = (((spread(quant(x, 6), 6, 2) << 1) | spread(quant(z, 6), 6, 2)) << 2) | quant(y, 2)
    elif mode == "polar":                 # azimuth 7 bits major, range 5 bits minor
        az = np.arctan2(x, z); r = np.hypot(x, z)
        m = (quant(az, 7) << 5) | quant(r, 5)
    elif mode == "polar_y":               # azimuth 6, range 4, height 2
        az = np.arctan2(x, z); r = np.hypot(x, z)
        m = (quant(az, 6) << 6) | (quant(r, 4) << 2) | quant(y, 2)
    else:
        raise ValueError(mode)
    return np.lexsort((np.arange(len(x)), m, cls))


def forms(K, H, W, xpose):
    ry, t = xpose[0], xpose[1:4]
    c, s = np.cos(ry), np.sin(ry)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    W1, H1 = W - 1.0, H - 1.0
    rows = [np.array([0, 0, 1.0]), np.array([fx, 0, cx]), np.array([fx, 0, cx - W1]), np.array([0, fy, cy]), np.array([0, fy, cy - H1])]
    return [(r @ R, float(r @ t)) for r in rows]          # coefficient vector on the raw point, constant


def pass_stats(p, lab, order, K, H, W, xpose):
    p = p[:, order]; lab = lab[order]
    n = p.shape[1]; G = (n + 31) // 32
    pad = G * 32 - n
    pp = np.concatenate([p, np.full((3, pad), np.nan)], 1).reshape(3, G, 32)
    ll = np.concatenate([lab, np.full(pad, -1)]).reshape(G, 32)
    lo = np.nanmin(pp, 2); hi = np.nanmax(pp, 2)
    ctr = 0.5 * (lo + hi); half = 0.5 * (hi - lo)
    f = forms(K, H, W, xpose)
    flo, fhi, val = [], [], []
    for a, c0 in f:
        mid = a @ ctr + c0; rad = np.abs(a) @ half
        flo.append(mid - rad); fhi.append(mid + rad)
        val.append(np.einsum("c,cgk->gk", a, np.nan_to_num(pp)) + c0)
    has0 = (ll == 0).any(1); has1 = (ll == 1).any(1)
    front = flo[0] > 0
    all_out = (fhi[0] < 0) | (front & (np.minimum(np.minimum(fhi[1], -flo[2]), np.minimum(fhi[3], -flo[4])) < 0))
    all_in = front & (np.minimum(np.minimum(flo[1], -fhi[2]), np.minimum(flo[3], -fhi[4])) > 0)
    skip = (~has0 | all_out) & (~has1 | all_in)
    sure = ~skip & ((has0 & ~has1 & all_in) | (has1 & ~has0 & all_out))
    und = ~skip & ~sure
    inside = (val[0] > 0) & (val[1] > 0) & (val[2] < 0) & (val[3] > 0) & (val[4] < 0)
    active = ((ll == 1) & ~inside) | ((ll == 0) & inside)
    return dict(groups=G, skip=int(skip.sum()), sure=int(sure.sum()), undecided=int(und.sum()), active=int(active.sum()),
                active_in_undecided=int(active[und].sum()), mixed=int((has0 & has1).sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=6)
    ap.add_argument("--inits", type=int, default=6)
    a = ap.parse_args()
    modes = ["none", "xz6", "xz7", "xz5y2", "xz6y2", "xyz4", "xyz5", "polar", "polar_y"]
    tot = {m: dict(skip=0, sure=0, undecided=0, active=0, groups=0, active_in_undecided=0, mixed=0) for m in modes}
    for s in range(a.samples):
        smp = syn.make_sample(100 + s)
        iy, pf, lf, _ = oracle.initial_guess(smp["points"], smp["pred"])
        pf = np.asarray(pf, dtype=np.float64); lf = np.asarray(lf)
        ry, t = syn.make_inits(100 + s, iy, a.inits)
        K = np.asarray(smp["K"], dtype=np.float64).reshape(3, 3)
        poses = [np.array([ry[i], t[i][0], t[i][1], t[i][2]]) for i in range(a.inits)]
        poses.append(np.array([smp["ry_gt"], *smp["t_gt"]]))          # around the solution, where most passes are spent
        for m in modes:
            order = keys(pf, lf, m)
            for xp in poses:
                st = pass_stats(pf, lf, order, K, smp["H"], smp["W"], xp)
                for k in tot[m]:
                    tot[m][k] += st[k]
    print("%-9s %8s %8s %8s %10s | %s" % ("order", "skip %", "sure %", "undec %", "act/pass", "modelled warp-instr / pass (box 70/round + 70/undecided group + 110/active batch)"))
    npass = a.samples * (a.inits + 1)
    for m in modes:
        v = tot[m]; G = v["groups"]
        und_g = v["undecided"] / npass; act = v["active"] / npass
        model = 70 * (G / npass / 32) + 70 * und_g + 110 * act / 32
        print("%-9s %8.1f %8.1f %8.1f %10.0f | %.0f   (mixed-label groups %.1f %%)" % (m, 100 * v["skip"] / G, 100 * v["sure"] / G, 100 * v["undecided"] / G, act, model, 100 * v["mixed"] / G))


if __name__ == "__main__":
    main()
