"""K0 on the device (tsdf_k0.cu) against the oracle's definition (oracle.cpp: k0_bilateral, k0_normals), bit for bit:
filtered depth and normals on clean, noisy and ragged frames; then the frame path with tsdf_config.preprocess = 1
(fusion, per-pixel linearisation, tracking) against the oracle with preprocess = 1, at the usual parity bars.
PCL itself is un-vendored: parity with PCL is unpinned (tests/test_k0.py states the definition)."""
import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tests.conftest import rot_angle
from tools import synth

pytestmark = pytest.mark.gpu


def bit_equal(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)


def cases(frames):
    depth, Rs, ts = frames
    yield "clean", depth[0]
    yield "noisy", synth.add_sensor_noise(depth[1], seed=1234)
    d = synth.add_sensor_noise(depth[2], seed=5, dropout=0.03)
    d[100:160, 200:330] = np.nan; d[300:310, 10:50] = 0.0; d[7, 9] = np.inf
    yield "ragged", d
    yield "empty", np.full_like(depth[0], np.nan)
    e = np.full_like(depth[0], np.nan); e[240, 320] = 1.0
    yield "one pixel", e
    yield "deep range", np.where(np.arange(640)[None, :] < 320, 0.5, 14.0).astype(np.float32) * np.ones((480, 1), np.float32)


def test_k0_bit_parity(gpu_lib, frames, K):
    g = T.Tsdf(T.default_config(m=32, preprocess=1)); g.set_intrinsics(K)
    o = po.Oracle(m=32, use_coord_table=0, preprocess=1); o.set_intrinsics(K)
    for name, d in cases(frames):
        zo, no = o.preprocess(d)
        zg, ng = g.preprocess(d)
        assert np.array_equal(zo, zg, equal_nan=True), (name, np.nanmax(np.abs(zo - zg)))
        assert np.array_equal(np.isnan(no), np.isnan(ng)), name
        assert np.array_equal(no, ng, equal_nan=True), (name, np.nanmax(np.abs(no - ng)))
    g.close(); o.close()


def test_frame_path_with_k0(gpu_lib, frames, K):
    depth, Rs, ts = frames
    noisy = synth.add_sensor_noise(depth[:6], seed=1234)
    kw = dict(m=64, preprocess=1, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(K)
    o = po.Oracle(use_coord_table=0, **kw); o.set_intrinsics(K)
    o.set_pose(Rs[0], ts[0])
    assert o.fuse(noisy[0]) == g.fuse(noisy[0], Rs[0], ts[0])
    for f in range(1, 5):
        Jo, po_, fo = o.linearize_pixels(noisy[f]); Jg, pg, fg = g.linearize_pixels(noisy[f])
        assert np.array_equal(fo, fg) and np.array_equal(Jo, Jg) and np.array_equal(po_, pg)
        st = o.track(noisy[f])
        Rg, tg, sg = g.track(noisy[f])
        Ro, to = o.get_pose()
        assert sg["n_valid"] == st["n_valid"] and np.abs(tg - to).max() < 1e-9 and rot_angle(Rg, Ro) < 1e-9
        assert o.fuse(noisy[f]) == g.fuse(noisy[f], Ro, to)
    Dg, Wg = g.download()
    assert np.array_equal(Dg, o.D) and np.array_equal(Wg, o.W)
    # K0 off: the same handle type without preprocessing sees different (noisier) normals
    g2 = T.Tsdf(T.default_config(m=64)); g2.set_intrinsics(K); g2.fuse(noisy[0], Rs[0], ts[0])
    assert not np.array_equal(g2.download()[0], Dg)
    with pytest.raises(T.TsdfError):
        g2.preprocess(noisy[0])
    g.close(); g2.close(); o.close()
