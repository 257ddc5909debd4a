#!/usr/bin/env python3
"""Absolute trajectory error (ATE) between an estimated and a ground-truth trajectory, TUM style
(the reference writes ./trajectory.txt for TUM's evaluate_ate.py,
sdf_reconstruction.cpp:4-17; the paper reports ATE RMSE on fr1/plant, Table I).

  python tools/evaluate_ate.py trajectory.txt data/fr1_plant_gt_every4.txt

Both files: `timestamp tx ty tz qx qy qz qw`.  Poses are associated by nearest timestamp (<= 0.02 s),
the estimate is rigidly aligned to the ground truth (Horn / Umeyama without scale) and the RMSE of
the translational residuals is printed.  ate_rmse() is also used by bench.py."""
import sys

import numpy as np


def align(est, gt):
    """Least-squares rigid alignment (rotation + translation) of est onto gt; both [n,3]."""
    mu_e, mu_g = est.mean(0), gt.mean(0)
    H = (est - mu_e).T @ (gt - mu_g)
    U, _, Vt = np.linalg.svd(H)
    S = np.eye(3)
    if np.linalg.det(Vt.T @ U.T) < 0:
        S[2, 2] = -1
    R = Vt.T @ S @ U.T
    t = mu_g - R @ mu_e
    return R, t


def ate_rmse(est_xyz, gt_xyz, do_align=True):
    est_xyz = np.asarray(est_xyz, float); gt_xyz = np.asarray(gt_xyz, float)
    if do_align:
        R, t = align(est_xyz, gt_xyz)
        est_xyz = est_xyz @ R.T + t
    err = np.linalg.norm(est_xyz - gt_xyz, axis=1)
    return float(np.sqrt((err ** 2).mean())), err


def associate(est, gt, max_dt=0.02):
    idx = np.searchsorted(gt[:, 0], est[:, 0])
    idx = np.clip(idx, 1, len(gt) - 1)
    left = gt[idx - 1, 0]; right = gt[idx, 0]
    idx = np.where(np.abs(est[:, 0] - left) < np.abs(est[:, 0] - right), idx - 1, idx)
    ok = np.abs(gt[idx, 0] - est[:, 0]) <= max_dt
    return np.nonzero(ok)[0], idx[ok]


def main():
    est = np.loadtxt(sys.argv[1], comments="#"); gt = np.loadtxt(sys.argv[2], comments="#")
    ie, ig = associate(est, gt)
    rmse, err = ate_rmse(est[ie, 1:4], gt[ig, 1:4])
    print("compared_pose_pairs %d" % len(ie))
    print("absolute_translational_error.rmse %.6f m" % rmse)
    print("absolute_translational_error.mean %.6f m" % err.mean())
    print("absolute_translational_error.max %.6f m" % err.max())


if __name__ == "__main__":
    main()
