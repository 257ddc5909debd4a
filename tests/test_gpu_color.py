"""GPU parity of colour fusion (sdf.cpp:294-304) and colour sampling (sdf.cpp:164-217) against the
CPU oracle, through the C ABI.  All of it is fp32/fp64 arithmetic restated operation for
operation, so the comparison is bit-exact."""
import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tests.conftest import rot_angle
from tools import synth

pytestmark = pytest.mark.gpu


def pair(m, K, **kw):
    o = po.Oracle(m=m, use_coord_table=0, metric=0, **kw); o.set_intrinsics(K)
    g = T.Tsdf(T.default_config(m=m, metric=0, **kw)); g.set_intrinsics(K)
    return o, g


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("m", [64, 96])
def test_color_fusion_bit_exact(gpu_lib, frames, K, m):
    depth, Rs, ts = frames
    o, g = pair(m, K)
    for f in range(4):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        o.set_pose(Rs[f], ts[f])
        n_o = o.fuse_rgb(depth[f], rgb)
        n_g = g.fuse_rgb(depth[f], rgb, Rs[f], ts[f])
        assert n_o == n_g and n_o > 0
    Dg, Wg = g.download(); cg = g.download_color()
    assert same(Dg, o.D) and same(Wg, o.W)
    co = o.color()
    for a, b, name in zip(cg, co, ("Color_W", "R", "G", "B")):
        assert same(a, b), name
    assert (co[0] > 0).sum() > 1000 and co[1][co[0] > 0].max() > 100
    # D/W are the same as without colour
    g2 = T.Tsdf(T.default_config(m=m, metric=0)); g2.set_intrinsics(K)
    for f in range(4):
        g2.fuse(depth[f], Rs[f], ts[f])
    D2, W2 = g2.download()
    assert same(D2, Dg) and same(W2, Wg)
    g.close(); g2.close(); o.close()


def test_color_with_ragged_validity_and_skewed_K(gpu_lib, frames, K):
    depth, Rs, ts = frames
    Ks = K.copy(); Ks[1] = 3.0                      # skew: certificates off, k_fuse_items colour path
    for Kc in (K, Ks):
        o, g = pair(64, Kc)
        for f in range(3):
            d = depth[f].copy()
            d[100:140, 200:420] = np.nan; d[::17, ::13] = np.nan; d[300:, :50] = 0.0
            rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
            o.set_pose(Rs[f], ts[f])
            assert o.fuse_rgb(d, rgb) == g.fuse_rgb(d, rgb, Rs[f], ts[f])
        for a, b in zip(g.download_color(), o.color()):
            assert same(a, b)
        g.close(); o.close()


def test_interpolate_color_parity(gpu_lib, frames, K):
    depth, Rs, ts = frames
    m = 64
    o, g = pair(m, K)
    for f in range(3):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        o.set_pose(Rs[f], ts[f]); o.fuse_rgb(depth[f], rgb); g.fuse_rgb(depth[f], rgb, Rs[f], ts[f])
    rng = np.random.default_rng(5)
    pts = rng.uniform([-3.2, -3.2, -0.7], [3.2, 3.2, 3.2], (20000, 3))
    # exact voxel centres of coloured voxels (the unscaled early return) and a few far outside
    cw = o.color()[0]
    ijk = np.argwhere(cw > 0)[:: max(1, int((cw > 0).sum() // 500))]
    centres = np.array([o.get_global_coordinates(q) for q in ijk[:500]])
    pts = np.concatenate([pts, centres, [[50, 50, 50], [-50, 0, 0]]])
    co = o.interpolate_color(pts); cg = g.interpolate_color(pts)
    assert same(co, cg)
    assert np.isfinite(co[:, 0]).sum() > 100 and np.isnan(co[-1, 0])
    g.close(); o.close()


def test_track_and_fuse_rgb(gpu_lib, frames, K):
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o, g = pair(64, K, **kw)
    g0 = T.Tsdf(T.default_config(m=64, metric=0, **kw)); g0.set_intrinsics(K)
    rgb0 = synth.synth_rgb(depth[0], Rs[0], ts[0])
    o.set_pose(Rs[0], ts[0]); o.fuse_rgb(depth[0], rgb0)
    g.fuse_rgb(depth[0], rgb0, Rs[0], ts[0]); g0.fuse(depth[0], Rs[0], ts[0])
    for f in range(1, 4):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        R, t, st, n = g.track_and_fuse_rgb(depth[f], rgb)
        R0, t0, st0, n0 = g0.track_and_fuse(depth[f])
        assert np.array_equal(R, R0) and np.array_equal(t, t0) and n == n0      # colour does not touch D/W or the pose
        o.track(depth[f]); Ro, to = o.get_pose()
        assert np.linalg.norm(t - to) < 1e-4 and rot_angle(R, Ro) < 1e-4
        o.set_pose(R, t)                                                        # identical poses -> identical fusion
        assert o.fuse_rgb(depth[f], rgb) == n
    for a, b in zip(g.download_color(), o.color()):
        assert same(a, b)
    g.close(); g0.close(); o.close()


def test_color_reset_and_errors(gpu_lib, frames, K):
    depth, Rs, ts = frames
    g = T.Tsdf(T.default_config(m=32, metric=0)); g.set_intrinsics(K)
    g.enable_color()
    cw, r, gg, b = g.download_color()
    assert (cw == 0).all() and (r == np.float32(0.4)).all() and (b == np.float32(0.4)).all()      # sdf.cpp:30-34
    rgb = synth.synth_rgb(depth[0], Rs[0], ts[0])
    g.fuse_rgb(depth[0], rgb, Rs[0], ts[0])
    assert (g.download_color()[0] > 0).any()
    g.reset()
    cw, r, gg, b = g.download_color()
    assert (cw == 0).all() and (gg == np.float32(0.4)).all()
    g.close()
    g1 = T.Tsdf(T.default_config(m=32, metric=1)); g1.set_intrinsics(K)
    with pytest.raises(T.TsdfError):
        g1.fuse_rgb(depth[0], rgb, Rs[0], ts[0])                                                  # no normal, no colour update in the reference
    g1.close()
