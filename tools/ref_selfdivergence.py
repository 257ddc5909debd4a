"""How reproducible is the REFERENCE ITSELF in closed loop?  BASELINE.json configs[0] (256^3, 100 frames, reference
defaults: 20 GN iterations, signed stop at 0.001) run free-running by oracle/_ref (the reference's own translation
units) with ONE OpenMP thread and with ALL host threads.  The only difference between the two runs is the order in
which the reference adds its per-thread normal-equation partial sums (camera_tracking.cpp:146-189); everything
downstream is the same code.  Prints the per-frame pose difference between the two runs and the grid difference at
the end — the yardstick for any free-running comparison against the reference (tests/test_gpu_config1.py).

    python tools/ref_selfdivergence.py [frames] [m]      (build container only: needs oracle/_ref)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po   # noqa: E402
from oracle import pyref as pr      # noqa: E402
from tools import synth             # noqa: E402


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra @ Rb.T) - 1.0) / 2.0, -1.0, 1.0)))


def run(n_threads, depth, Rs, ts, m):
    pr.set_num_threads(n_threads); po.set_num_threads(n_threads)
    r = pr.Reference(m=m)
    r.set_intrinsics(synth.K_DEFAULT)
    r.set_pose(Rs[0], ts[0]); r.fuse(depth[0], count=False)
    poses, its = [], []
    for f in range(1, len(depth)):
        st = r.track(depth[f]); r.fuse(depth[f], count=False)
        R, t = r.get_pose()
        poses.append((R.copy(), t.copy())); its.append(st["iterations"])
    D, W = r.D.copy(), r.W.copy()
    r.close()
    return poses, its, D, W


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    depth, Rs, ts = synth.render_sequence(nf)
    nthr = len(os.sched_getaffinity(0))
    p1, i1, D1, W1 = run(1, depth, Rs, ts, m)
    pn, in_, Dn, Wn = run(nthr, depth, Rs, ts, m)
    dt = [float(np.linalg.norm(a[1] - b[1])) for a, b in zip(p1, pn)]
    dr = [rot_angle(a[0], b[0]) for a, b in zip(p1, pn)]
    first = next((f + 1 for f, (a, b) in enumerate(zip(i1, in_)) if a != b), None)
    rep = {"what": "reference vs reference: 1 OpenMP thread vs %d threads, free-running" % nthr, "frames": nf, "m": m,
           "pose_diff_m": {"max": max(dt), "median": float(np.median(dt)), "first_frame_above_1e-9": next((f + 1 for f, x in enumerate(dt) if x > 1e-9), None),
                           "first_frame_above_1e-4": next((f + 1 for f, x in enumerate(dt) if x > 1e-4), None)},
           "pose_diff_rad": {"max": max(dr), "median": float(np.median(dr))},
           "first_frame_with_different_iteration_count": first,
           "frac_dD_gt_1e-6": float((np.abs(D1 - Dn) > 1e-6).mean()), "frac_dW_gt_1e-6": float((np.abs(W1 - Wn) > 1e-6).mean()),
           "pose_diff_m_per_frame": dt, "pose_diff_rad_per_frame": dr}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
