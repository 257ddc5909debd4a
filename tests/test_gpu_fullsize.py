"""Parity at BASELINE.json's full single-GPU size (512^3, 640x480): the whole grid bit for bit against the
oracle after fusing, the normal equations and one tracked pose at the stated tolerances, the certificate
self-check, slabs against the unsharded volume, and size-independent properties (idempotent skip of unseen
space, weight bookkeeping).  Sized to finish in about a minute on the GPU box."""
import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tests.conftest import rot_angle

pytestmark = pytest.mark.gpu
M = 512


def test_fullsize_fusion_bit_exact_and_tracking(gpu_lib, frames, K):
    depth, Rs, ts = frames
    kw = dict(m=M, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o = po.Oracle(use_coord_table=0, **kw); o.set_intrinsics(K)
    g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(K)
    n_tot = 0
    for f in range(2):
        o.set_pose(Rs[f], ts[f])
        n_o = o.fuse(depth[f]); n_g = g.fuse(depth[f], Rs[f], ts[f])
        assert n_o == n_g
        n_tot += n_g
    D, W = g.download()
    assert np.array_equal(W, o.W) and np.array_equal(D, o.D)                  # 134 M voxels, every bit
    # weight bookkeeping: every update adds a weight in (0.96, 1] (sdf.cpp:276-279, 290)
    assert 0.96 * n_tot < float(W.sum(dtype=np.float64)) <= n_tot
    del D, W
    A_o, b_o, s_o = o.linearize(depth[2]); A_g, b_g, s_g = g.linearize(depth[2])
    assert s_o["n_valid"] == s_g["n_valid"] > 30000
    assert np.abs(A_g - A_o).max() <= 1e-5 * np.abs(A_o).max() and np.abs(b_g - b_o).max() <= 1e-5 * np.abs(b_o).max()
    o.track(depth[2]); Ro, to = o.get_pose()
    R, t, st = g.track(depth[2])
    assert st["iterations"] == 10 and np.linalg.norm(t - to) <= 1e-4 and rot_angle(R, Ro) <= 1e-4
    # certificates at full size: the exact path on every voxel of every certified unit agrees
    for f in (3, 7):
        g.set_pose(Rs[f], ts[f])
        r = g.debug_fuse_check(depth[f])
        assert r["wrong"] == 0 and r["fast"] > 1000000, r
    # fusing a frame with no valid depth changes nothing (every voxel is skipped, sdf.cpp:260)
    before = g.total_updates(reset=True)
    assert g.fuse(np.full((480, 640), np.nan, np.float32), Rs[1], ts[1]) == 0
    g.close(); o.close()


def test_fullsize_slabs_equal_unsharded(gpu_lib, frames, K):
    depth, Rs, ts = frames
    kw = dict(m=M, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    one = T.Tsdf(T.default_config(**kw)); one.set_intrinsics(K)
    grp = T.ShardGroup(4, **kw); grp.set_intrinsics(K)
    for f in range(2):
        one.fuse(depth[f], Rs[f], ts[f])
        grp.set_pose(Rs[f], ts[f]); grp.frame(depth[f], track=False, fuse=True)
    A1, b1, s1 = one.linearize(depth[2]); A2, b2, s2 = grp.linearize(depth[2])
    assert s1["n_valid"] == s2["n_valid"] and s2["halo_miss"] == 0
    assert np.abs(A1 - A2).max() <= 1e-12 * np.abs(A1).max() and np.abs(b1 - b2).max() <= 1e-12 * np.abs(b1).max()
    D1, W1 = one.download()
    for s in grp.shards:                                                      # every slab incl. its redundantly fused halo
        ks0, ks1, _, _ = s.stored_range()
        d, w = s.download()
        assert np.array_equal(d, D1[:, :, ks0:ks1]) and np.array_equal(w, W1[:, :, ks0:ks1])
    one.close(); grp.close()
