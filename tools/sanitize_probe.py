"""Small end-to-end pass over every kernel family (64^3) for compute-sanitizer:
prep, pyramid, linearize (+ GN update), fusion (tables, plan, cert with affine end points, exact, items with skewed K),
colour fusion (incl. the dense pose: rows certified free straight from the item list), colour sampling, the one-sweep
mesher (list emit, and the two-sweep fallback on a list overflow), K0 (bilateral grid + window normals), accessors,
in-process z-slab shards (micro-tile sweeps).  Run as
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
# K0 on noisy / ragged depth, through the frame path
gk = T.Tsdf(T.default_config(preprocess=1, **kw)); gk.set_intrinsics(K)
noisy = synth.add_sensor_noise(depth, seed=3, dropout=0.02)
gk.preprocess(noisy[0]); gk.fuse(noisy[0], Rs[0], ts[0]); gk.track_and_fuse(noisy[1])
gk.preprocess(np.full_like(depth[0], np.nan))
gk.close()
# dense pose with colour: every row certified free -> exact<colour> consumes items directly; large mesh -> list overflow fallback
gd = T.Tsdf(T.default_config(m=128)); gd.set_intrinsics(K); gd.enable_color()
Rd = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float); td = np.array([0.0, -12.0, 1.25])
gd.fuse_rgb(np.full((480, 640), 40.0, np.float32), np.full((480, 640, 3), 128, np.uint8), Rd, td)
gd.close()
grp = T.ShardGroup(2, **kw); grp.set_intrinsics(K); grp.set_pose(Rs[0], ts[0])
grp.frame(depth[0], track=False, fuse=True); grp.frame(depth[1], track=True, fuse=True)
grp.close()
print("sanitize probe done: %d mesh vertices, %d voxels updated" % (len(xyz), n))
