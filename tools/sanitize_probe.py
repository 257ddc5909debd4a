"""Small end-to-end pass over every kernel family (64^3) for compute-sanitizer:
prep, pyramid, linearize (+ GN update), fusion (tables, plan, cert, exact, items with skewed K), colour fusion,
colour sampling, mesher, accessors, in-process z-slab shards.  Run as
    compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_probe.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth

depth, Rs, ts = synth.render_sequence(4)
K = synth.K_DEFAULT
kw = dict(m=64, gauss_newton_max_iteration=3, maximum_twist_diff=float("-inf"))
g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(K)
g.fuse(depth[0], Rs[0], ts[0])
R, t, st, n = g.track_and_fuse(depth[1])
d = depth[2].copy(); d[100:200, 100:300] = np.nan
g.track(d)
g.fuse_rgb(depth[2], synth.synth_rgb(depth[2], Rs[2], ts[2]), Rs[2], ts[2])
xyz, world, rgba = g.mesh(0.0, world=True, colors=True)
g.interpolate_distance(np.random.default_rng(0).uniform(-1, 65, (256, 3)))
g.interpolate_color(world[:64])
D, W = g.download(); g.upload(D, W); g.download_color()
g.backproject(depth[0])
for f in range(3):                       # streaming paths
    g.submit_frame(depth[f], track=1, slot=f)
g.sync()
g.close()
Ks = K.copy(); Ks[1] = 2.0
g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(Ks)
g.fuse(depth[0], Rs[0], ts[0]); g.fuse_rgb(depth[1], synth.synth_rgb(depth[1], Rs[1], ts[1]), Rs[1], ts[1])
g.close()
grp = T.ShardGroup(2, **kw); grp.set_intrinsics(K); grp.set_pose(Rs[0], ts[0])
grp.frame(depth[0], track=False, fuse=True); grp.frame(depth[1], track=True, fuse=True)
grp.close()
print("sanitize probe done: %d mesh vertices, %d voxels updated" % (len(xyz), n))
