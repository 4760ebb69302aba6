"""Hand-off between the classifier and the registration solver.

The reference passes data through a triple of files per frame, written by
evaluation/visualize_and_save_data.py:174-186 and read back by evaluation/registration_lsq.py:291-302:

    <id>_pc_label.npy   float array [7, N]: rows 0-2 xyz, 3 coarse prediction, 4 coarse label,
                        5 fine prediction, 6 fine label (labels stored as floats)
    <id>_K.npy          [3, 3] intrinsics
    <id>_P.npy          [3, 4] or [4, 4] ground-truth pose

with <id> = '%06d_%02d' (9 characters, registration_lsq.py:273).  The batched contract of this
framework is the C ABI's device record: (xyz f32 [S,3,Ns], label int8 [S,Ns], n_pts int32 [S],
K f64 [S,9]).  This module converts between the two on the host (file I/O only -- no compute path
lives here) and drives a whole directory through frustum.register_batch.
"""
import math
import os

import numpy as np

ROW_COARSE_PREDICTION = 3
ROW_COARSE_LABEL = 4
ROW_FINE_PREDICTION = 5
ROW_FINE_LABEL = 6
LABEL_ROWS = {"coarse_prediction": ROW_COARSE_PREDICTION, "coarse_label": ROW_COARSE_LABEL,
              "fine_prediction": ROW_FINE_PREDICTION, "fine_label": ROW_FINE_LABEL}
SUFFIXES = ("_pc_label.npy", "_K.npy", "_P.npy")

# registration_lsq.py:236-247: nuScenes clouds are east-north-up; camera axes are x right, y down, z forward.
ENU2CAM = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


def list_records(data_dir):
    """Frame ids present in a legacy directory: first 9 characters of every file name, unique, sorted
    (registration_lsq.py:273-275; the reference then shuffles them, which only changes print order)."""
    names = {f[0:9] for f in os.listdir(data_dir) if os.path.isfile(os.path.join(data_dir, f)) and f.endswith(".npy")
             and f.endswith(SUFFIXES)}
    return sorted(names)


def save_record(data_dir, name, pc, coarse_prediction, coarse_label, fine_prediction, fine_label, K, P):
    """Write one frame in the reference's layout (visualize_and_save_data.py:174-186)."""
    pc = np.asarray(pc, dtype=np.float32)
    rows = [np.asarray(r, dtype=np.float32)[None, :] for r in (coarse_prediction, coarse_label, fine_prediction,
                                                               fine_label)]
    out = np.concatenate([pc] + rows, axis=0)
    os.makedirs(data_dir, exist_ok=True)
    np.save(os.path.join(data_dir, name + "_pc_label.npy"), out)
    np.save(os.path.join(data_dir, name + "_K.npy"), np.asarray(K))
    np.save(os.path.join(data_dir, name + "_P.npy"), np.asarray(P))


def load_record(data_dir, name, which="coarse_prediction", enu2cam=False):
    """One frame -> (pc [3,N] float, label [N] int64, K [3,3] f64, P_gt [4,4] f64), as
    registration_lsq.py:291-304 reads it."""
    data = np.load(os.path.join(data_dir, name + "_pc_label.npy"))
    if data.ndim != 2 or data.shape[0] < 4:
        raise ValueError("%s_pc_label.npy: expected [>=4, N], got %s" % (name, data.shape))
    row = LABEL_ROWS[which]
    if row >= data.shape[0]:
        raise ValueError("%s_pc_label.npy has no row %d (%s)" % (name, row, which))
    pc = data[0:3, :]
    label = data[row, :].astype(np.int64)
    K = np.load(os.path.join(data_dir, name + "_K.npy")).astype(np.float64)
    P = np.load(os.path.join(data_dir, name + "_P.npy")).astype(np.float64)
    if P.shape[0] == 3:
        P = np.concatenate((P, np.identity(4)[3:4, :]), axis=0)
    if enu2cam:
        pc = (ENU2CAM[0:3, 0:3].astype(pc.dtype) @ pc)
        P = P @ np.linalg.inv(ENU2CAM)
    return pc, label, K, P


def load_batch(data_dir, names=None, which="coarse_prediction", enu2cam=False):
    """Legacy directory -> host arrays of the batched contract.

    Returns dict(names, xyz f32 [S,3,Nmax], label int8 [S,Nmax] (-1 = padding), n_pts int32 [S],
    K f64 [S,9], P_gt f64 [S,4,4]).  Coordinates are float32: that is what the loaders produced
    before the .npy detour promoted them (data/kitti_pc_img_pose_loader.py:431)."""
    if names is None:
        names = list_records(data_dir)
    recs = [load_record(data_dir, n, which, enu2cam) for n in names]
    S = len(recs)
    n_max = max((r[0].shape[1] for r in recs), default=0)
    xyz = np.zeros((S, 3, n_max), dtype=np.float32)
    label = np.full((S, n_max), -1, dtype=np.int8)
    n_pts = np.zeros((S,), dtype=np.int32)
    K = np.zeros((S, 9), dtype=np.float64)
    P = np.zeros((S, 4, 4), dtype=np.float64)
    for s, (pc, lab, k, p) in enumerate(recs):
        n = pc.shape[1]
        pc32 = pc.astype(np.float32)
        if pc.dtype != np.float32 and not np.array_equal(pc32.astype(pc.dtype), pc):
            raise ValueError("%s: coordinates are not float32-representable; use frustum.pack_clouds" % names[s])
        xyz[s, :, :n] = pc32
        label[s, :n] = np.where(lab == 1, 1, np.where(lab == 0, 0, -1)).astype(np.int8)
        n_pts[s] = n
        K[s] = k.reshape(9)
        P[s] = p
    return dict(names=list(names), xyz=xyz, label=label, n_pts=n_pts, K=K, P_gt=P)


def summarize(t_err, r_err, cost, ok):
    """registration_result_analysis.py:19-47 on the per-frame errors: frames with cost <= 1e-6 are dropped,
    RTE/RRE mean and sigma, success rate over the kept frames."""
    t_err, r_err, cost, ok = (np.asarray(a) for a in (t_err, r_err, cost, ok))
    valid = cost > 1e-6
    t, r, s = t_err[valid], r_err[valid], ok[valid]
    n = int(valid.sum())
    if n == 0:
        return dict(n=0, rte_mean=math.nan, rte_sigma=math.nan, rre_mean=math.nan, rre_sigma=math.nan,
                    success_rate=math.nan)
    return dict(n=n, rte_mean=float(t.mean()), rte_sigma=float(math.sqrt(t.var())), rre_mean=float(r.mean()),
                rre_sigma=float(math.sqrt(r.var())), success_rate=float(s.astype(np.float64).mean()))


def register_directory(data_dir, H, W, which="coarse_prediction", enu2cam=False, n_inits=60, seed=0, is_2d=True,
                       batch=512, names=None, device="cuda", out_dir=None):
    """The __main__ of evaluation/registration_lsq.py:250-398 as one function: every frame of a legacy
    directory through frustum.register_batch (the GPU path), errors by frustum.pose_error_batch, and the
    P_pred_all_np / P_gt_all_np / cost_all_np files the analysis script expects (:396-398)."""
    import torch
    from . import frustum

    if names is None:
        names = list_records(data_dir)
    P_pred = np.zeros((len(names), 4, 4))
    P_gt = np.zeros((len(names), 4, 4))
    cost = np.zeros((len(names),))
    for a in range(0, len(names), batch):
        rec = load_batch(data_dir, names[a:a + batch], which, enu2cam)
        n_in = rec["xyz"].shape[2]
        ns = frustum.round_up(max(n_in, 1), 16)
        xyz = torch.zeros((len(rec["names"]), 3, ns), dtype=torch.float32, device=device)
        pred = torch.full((len(rec["names"]), ns), -1, dtype=torch.int8, device=device)
        xyz[:, :, :n_in] = torch.from_numpy(rec["xyz"]).to(device)
        pred[:, :n_in] = torch.from_numpy(rec["label"]).to(device)
        out = frustum.register_batch(xyz, pred, n_in, rec["K"], H, W, n_inits=n_inits, seed=seed + a, is_2d=is_2d)
        P_pred[a:a + batch] = out["P"].cpu().numpy()
        cost[a:a + batch] = out["cost"].cpu().numpy()
        P_gt[a:a + batch] = rec["P_gt"]
    err = frustum.pose_error_batch(P_pred, P_gt)
    t_err, r_err, ok = err["t_err"].cpu().numpy(), err["r_err"].cpu().numpy(), err["success"].cpu().numpy()
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "P_pred_all_np.npy"), P_pred)
        np.save(os.path.join(out_dir, "P_gt_all_np.npy"), P_gt)
        np.save(os.path.join(out_dir, "cost_all_np.npy"), cost)
    return dict(names=names, P_pred=P_pred, P_gt=P_gt, cost=cost, t_err=t_err, r_err=r_err, success=ok,
                summary=summarize(t_err, r_err, cost, ok))
