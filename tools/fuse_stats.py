"""Certificate statistics + per-kernel timings of fusion at 512^3 (run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth, Rs, ts = synth.render_sequence(40)
g = T.Tsdf(T.default_config(m=m)); g.set_intrinsics(synth.K_DEFAULT)
for f in (0, 10, 20, 39):
    g.set_pose(Rs[f], ts[f])
    r = g.debug_fuse_check(depth[f])
    n = g.fuse(depth[f])
    print("frame %d: items %d (%.1fM voxel slots)  certified voxels %.2fM  wrong %d  updated %.2fM" % (f, r["items"], r["items"] * 128 / 1e6, r["fast"] / 1e6, r["wrong"], n / 1e6))
