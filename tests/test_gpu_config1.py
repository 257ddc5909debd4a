"""BASELINE.json configs[0] as written: 256^3 grid, 640x480 synthetic depth, 100 frames along the bundled
fr1/plant ground-truth path, the reference's own parameters (20 GN iterations, signed stop test at 0.001,
sdf_reconstruction.cpp:83-88) — the GPU path FREE-RUNNING (tsdf_track_and_fuse, no resynchronisation) against
the CPU oracle free-running on the same frames.

What can differ: the double-precision sums of the normal equations are taken in a different (fixed) order on the
GPU, so poses differ at ~1e-15 per frame; the reference's discontinuous arithmetic ((int) truncation of pixel and
voxel coordinates, W > 0, d > delta, the signed stop test) turns such a difference into different decisions for
single voxels / pixels, and the closed loop amplifies them.

THE YARDSTICK IS THE REFERENCE ITSELF: profiles/r02_ref_selfdivergence.json (tools/ref_selfdivergence.py) runs this
very configuration with oracle/_ref — the reference's own translation units — once with 1 OpenMP thread and once
with 8, which only changes the order of ITS per-thread partial sums (camera_tracking.cpp:146-189).  The two runs of
the reference agree to < 1e-9 m for 14 frames, pass 1e-4 m at frame 29, reach 1.3 mm / 5.0e-3 rad, run different
iteration counts from frame 58 on and end with 8 % of the voxels more than 1e-6 apart.  A free-running comparison
over 100 frames therefore cannot be held to north_star's per-step tolerances (those are asserted step-wise, same
state in -> one step -> compare, in test_gpu_parity.py); what this test asserts is
  * 1e-4 m / 1e-4 rad on the first 15 tracked frames (before the loop has amplified anything),
  * over all 100 frames, a divergence no larger than the reference's own (with head room), and
  * that the path tracks (error against the ground-truth path);
and it REPORTS everything else — per-frame maxima, iteration-count agreement, the |dD| / |dW| histogram — to
gpurun_out/config1_parity.json (kept under profiles/).
"""
import json
import os

import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tests.conftest import rot_angle
from tools import synth

pytestmark = pytest.mark.gpu

N_FRAMES = 100
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config1_closed_loop_100_frames(gpu_lib):
    depth, Rs, ts = synth.render_sequence(N_FRAMES)
    g = T.Tsdf(T.default_config(m=256)); g.set_intrinsics(synth.K_DEFAULT)          # reference defaults: 20 its, 0.001
    o = po.Oracle(m=256, use_coord_table=0); o.set_intrinsics(synth.K_DEFAULT)
    o.set_pose(Rs[0], ts[0]); o.fuse(depth[0]); g.fuse(depth[0], Rs[0], ts[0])
    dt, dr, its_equal, nupd_rel, gt_err = [], [], 0, [], []
    for f in range(1, N_FRAMES):
        st = o.track(depth[f]); n_o = o.fuse(depth[f])
        Rg, tg, sg, n_g = g.track_and_fuse(depth[f])
        Ro, to = o.get_pose()
        dt.append(float(np.linalg.norm(tg - to))); dr.append(rot_angle(Rg, Ro))
        its_equal += int(sg["iterations"] == st["iterations"] and sg["stopped"] == st["stopped"])
        nupd_rel.append(abs(n_g - n_o) / max(n_o, 1))
        gt_err.append(float(np.linalg.norm(tg - ts[f])))
        assert st["n_oob"] == 0 and sg["n_oob"] == 0                                 # TRAP 5 never fires on the benchmark inputs
        if f <= 15:
            assert dt[-1] <= 1e-4 and dr[-1] <= 1e-4, (f, dt[-1], dr[-1])            # north_star: 1e-4 m / 1e-4 rad
        assert dt[-1] <= 5e-3 and dr[-1] <= 2e-2, (f, dt[-1], dr[-1])                # reference vs itself: 1.3e-3 m / 5.0e-3 rad
    Dg, Wg = g.download()
    dD = np.abs(Dg - o.D); dW = np.abs(Wg - o.W)
    edges = [0.0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, np.inf]
    seen = (o.W > 0) | (Wg > 0)
    rep = {"frames": N_FRAMES, "m": 256, "gn": "reference defaults (20 iterations, signed stop at 0.001)",
           "pose_diff_m_max": max(dt), "pose_diff_rad_max": max(dr), "pose_diff_m_median": float(np.median(dt)),
           "first_frame_above_1e-9_m": next((q + 1 for q, x in enumerate(dt) if x > 1e-9), None),
           "first_frame_above_1e-4_m": next((q + 1 for q, x in enumerate(dt) if x > 1e-4), None),
           "yardstick": "profiles/r02_ref_selfdivergence.json: the reference against itself (1 vs 8 threads): 1.3e-3 m, 5.0e-3 rad, "
                        "1e-4 m passed at frame 29, 8.0 % of voxels > 1e-6",
           "frames_with_equal_iteration_count": its_equal, "tracked_frames": N_FRAMES - 1,
           "n_updated_rel_diff_max": max(nupd_rel), "tracking_err_vs_gt_m_final": gt_err[-1], "tracking_err_vs_gt_m_max": max(gt_err),
           "voxels": int(Dg.size), "voxels_seen": int(seen.sum()),
           "voxels_bit_equal": int(((Dg == o.D) & (Wg == o.W)).sum()),
           "frac_dD_gt_1e-6": float((dD > 1e-6).mean()), "frac_dW_gt_1e-6": float((dW > 1e-6).mean()),
           "hist_edges": [e if np.isfinite(e) else "inf" for e in edges],
           "hist_dD": np.histogram(dD, bins=edges)[0].tolist(), "hist_dW": np.histogram(dW, bins=edges)[0].tolist(),
           "max_dD": float(dD.max()), "max_dW": float(dW.max())}
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        json.dump(rep, open(os.path.join(out_dir, "config1_parity.json"), "w"), indent=1)
    except OSError:
        pass
    print("config1:", json.dumps(rep))
    assert max(gt_err) < 0.10                             # the path itself tracks (2.3 cm voxels)
    # no worse than the reference against itself (8.0 % / 5.7 % of voxels beyond 1e-6), with head room
    assert rep["frac_dD_gt_1e-6"] <= 0.16 and rep["frac_dW_gt_1e-6"] <= 0.12, rep
    g.close(); o.close()
