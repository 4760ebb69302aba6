#!/usr/bin/env python
"""Register every frame of a legacy hand-off directory on the GPU.

Replaces `python evaluation/registration_lsq.py` + `registration_result_analysis.py` of the reference
(registration_lsq.py:250-398, registration_result_analysis.py:13-47) for directories written by
visualize_and_save_data.py:174-186.  Example:

    python scripts/register_dir.py /data/kitti/save/run/data --H 160 --W 512 --out /data/kitti/save/run
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from deepi2p_b200 import handoff  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("data_dir")
    ap.add_argument("--H", type=float, required=True, help="image height (kitti 160, oxford 384, nuscenes 160)")
    ap.add_argument("--W", type=float, required=True, help="image width (kitti 512, oxford 640, nuscenes 320)")
    ap.add_argument("--labels", default="coarse_prediction", choices=sorted(handoff.LABEL_ROWS))
    ap.add_argument("--enu2cam", action="store_true", help="nuScenes axis convention (registration_lsq.py:236-247)")
    ap.add_argument("--inits", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--is-3d", action="store_true")
    ap.add_argument("--every", type=int, default=1, help="take every n-th frame (the reference uses 30, :285)")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--out", default=None, help="where to write P_pred_all_np.npy / P_gt_all_np.npy / cost_all_np.npy")
    args = ap.parse_args()

    names = handoff.list_records(args.data_dir)[::args.every]
    res = handoff.register_directory(args.data_dir, args.H, args.W, which=args.labels, enu2cam=args.enu2cam,
                                     n_inits=args.inits, seed=args.seed, is_2d=not args.is_3d, batch=args.batch,
                                     names=names, out_dir=args.out)
    for i, n in enumerate(res["names"]):
        print("%s - cost: %.1f, T: %.1f, R:%.1f" % (n, res["cost"][i], res["t_err"][i], res["r_err"][i]))
    s = res["summary"]
    print("RTE %.2f +- %.2f, RRE %.2f +- %.2f, success rate %.2f" % (s["rte_mean"], s["rte_sigma"], s["rre_mean"],
                                                                     s["rre_sigma"], s["success_rate"] * 100))
    print(json.dumps(s))


if __name__ == "__main__":
    main()
