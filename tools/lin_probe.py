"""Phase timing of the linearise kernel at 512^3 (debug aid, run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth, Rs, ts = synth.render_sequence(8)
g = T.Tsdf(T.default_config(m=m, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf")))
g.set_intrinsics(synth.K_DEFAULT); g.set_pose(Rs[0], ts[0]); g.fuse(depth[0])
import os
for f in range(1, 6):
    if os.environ.get("LIN_PROBE_GT"): g.fuse(depth[f], Rs[f], ts[f])      # experiment builds that cannot track
    else: g.track_and_fuse(depth[f])
for rep in range(4):
    t = g.debug_phase_times(depth[6])
    print("main loop %.2f us | to final block %.2f | reduce %.2f | update %.2f | total %.2f" % tuple(
        (b - a) / 1e3 for a, b in [(t[0], t[1]), (t[1], t[2]), (t[2], t[3]), (t[3], t[4]), (t[0], t[4])]))
