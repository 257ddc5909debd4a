import sys; sys.path.insert(0, ".")
import numpy as np
import tracking_sdf_b200 as T
from tools import synth
depth, Rs, ts = synth.render_sequence(6)
K = synth.K_DEFAULT
n_shards, m = 4, 128
kw = dict(m=m, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
grp = T.ShardGroup(n_shards, **kw); grp.set_intrinsics(K)
grp.set_pose(Rs[0], ts[0])
print(grp.frame(depth[0], track=False, fuse=True)[3])
for f in range(1, 4):
    grp.linearize(depth[f]); print("lin ok", f)
    grp.frame(depth[f], track=True, fuse=True); print("frame ok", f)
for s in grp.shards: s.reset()
for f in range(3):
    grp.set_pose(Rs[f], ts[f]); grp.frame(depth[f], track=False, fuse=True); print("fuse ok", f)
