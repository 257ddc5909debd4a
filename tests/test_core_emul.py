"""CPU check of the CUDA kernels' __host__ __device__ core (tracking_sdf_b200/csrc/tsdf_core.cuh,
compiled for the HOST by tests/host_emul) against the oracle — bit for bit.

This is NOT the GPU parity test (those are tests/test_gpu_*.py, -m gpu, through the C ABI); it
exists because the build container has no GPU, so the kernels' exact-arithmetic building
blocks (hoisted camera-space sums, scan-line clip, sample interpolation, J/psi, GN update) are
verified here first.  The emulation library is test infrastructure and never loaded by the product.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import emul
from tools import synth


def _pair(m, metric, K, **kw):
    o = po.Oracle(m=m, use_coord_table=0, metric=metric, **kw)
    o.set_intrinsics(K)
    e = emul.Emul(K, m=m, metric=metric,
                  max_twist_diff=kw.get("maximum_twist_diff", 0.001), max_iter=kw.get("gauss_newton_max_iteration", 20))
    return o, e


@pytest.mark.parametrize("metric", [0, 1])
def test_prep_pose_fuse_linearize_bit_exact(metric, frames, K):
    depth, Rs, ts = frames
    o, e = _pair(64, metric, K)
    o.set_pose(Rs[0], ts[0]); e.set_pose(Rs[0], ts[0])
    assert np.array_equal(o.get_pose_inv()[0], e.get_pose()[2]) and np.array_equal(o.get_pose_inv()[1], e.get_pose()[3])
    co, no = o.backproject(depth[0]); e.prep(depth[0]); ce, ne = e.cloud()
    assert np.array_equal(co, ce, equal_nan=True) and np.array_equal(no, ne, equal_nan=True)
    for f in range(3):
        o.set_pose(Rs[f], ts[f]); e.set_pose(Rs[f], ts[f]); e.prep(depth[f])
        assert o.fuse(depth[f]) == e.fuse(use_clip=1)
        assert np.array_equal(o.D, e.D) and np.array_equal(o.W, e.W)
    o.set_pose(Rs[3], ts[3]); e.set_pose(Rs[3], ts[3]); e.prep(depth[3])
    Jo, po_, fo = o.linearize_pixels(depth[3]); Je, pe, fe, sums = e.linearize()
    assert np.array_equal(fo, fe) and np.array_equal(Jo, Je) and np.array_equal(po_, pe)
    assert (fo == 1).sum() > 30000
    A, b, st = o.linearize(depth[3]); Ae, be = emul.sums_to_Ab(sums)
    assert np.abs(A - Ae).max() <= 1e-12 * np.abs(A).max() and np.abs(b - be).max() <= 1e-12 * np.abs(b).max()
    assert sums[28] == st["n_valid"] and sums[29] == st["n_oob"] and sums[27] == pytest.approx(st["residual"], rel=1e-12)
    tw_o, sing = o.apply_update(A, b); stt, tw_e = e.gn_update(sums)
    assert np.abs(tw_o - tw_e).max() < 1e-12 and not sing and not stt["singular"]
    assert np.abs(o.get_pose()[0] - e.get_pose()[0]).max() < 1e-12 and np.abs(o.get_pose()[1] - e.get_pose()[1]).max() < 1e-12
    o.close()


def test_scanline_clip_never_drops_a_voxel(frames, K):
    # the clip is a conservative superset: fusing with and without it must give the same grid
    depth, Rs, ts = frames
    for f, m in [(0, 48), (5, 64), (9, 32)]:
        a = emul.Emul(K, m=m); b = emul.Emul(K, m=m)
        for e in (a, b):
            e.set_pose(Rs[f], ts[f]); e.prep(depth[f])
        na = a.fuse(use_clip=1); nb = b.fuse(use_clip=0)
        assert na == nb and np.array_equal(a.grid, b.grid)
    # camera outside the volume looking in (the dense micro-benchmark pose): every voxel is updated
    e = emul.Emul(K, m=32)
    R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float)      # optical axis along world +y
    e.set_pose(R, [0.0, -12.0, 1.25])
    e.prep(np.full((480, 640), 40.0, np.float32))
    assert e.fuse(use_clip=1) == 32 ** 3
    assert (e.W == 1).all() and (e.D == np.float32(-0.3)).all()


def test_certificates_match_exact_path(frames, K):
    # the pyramid certificates (free space / skipped) must never change a result: fuse with and without them
    depth, Rs, ts = frames
    for metric in (0, 1):
        for f, m in [(0, 64), (4, 96), (8, 48)]:
            a = emul.Emul(K, m=m, metric=metric); b = emul.Emul(K, m=m, metric=metric)
            for e in (a, b):
                e.set_pose(Rs[f], ts[f]); e.prep(depth[f])
            na = a.fuse(use_fast=1); nb = b.fuse(use_fast=0)
            assert na == nb and np.array_equal(a.grid, b.grid)
            assert a.last_fast > 0.3 * na                    # and they decide a good share of the in-view voxels
            g = (f + 1) % len(depth)                         # second frame on a non-empty grid
            for e in (a, b):
                e.set_pose(Rs[g], ts[g]); e.prep(depth[g])
            assert a.fuse(use_fast=1) == b.fuse(use_fast=0) and np.array_equal(a.grid, b.grid)
    # ragged validity: invalid pixels must make units skip or fall back, never update
    d = depth[2].copy(); d[100:300, 200:500] = np.nan; d[::7, ::5] = 0.0
    a = emul.Emul(K, m=64); b = emul.Emul(K, m=64)
    for e in (a, b):
        e.set_pose(Rs[2], ts[2]); e.prep(d)
    assert a.fuse(use_fast=1) == b.fuse(use_fast=0) and np.array_equal(a.grid, b.grid)


def test_skewed_intrinsics_take_the_general_projection(frames):
    # K with skew: camera_tracking.cpp:44 keeps the full 3x3 product
    depth, Rs, ts = frames
    Ks = synth.K_DEFAULT.copy(); Ks[1] = 0.7
    o, e = _pair(48, 1, Ks)
    o.set_pose(Rs[2], ts[2]); e.set_pose(Rs[2], ts[2]); e.prep(depth[2])
    assert o.fuse(depth[2]) == e.fuse(1)
    assert np.array_equal(o.D, e.D) and np.array_equal(o.W, e.W)
    o.close()


def test_invalid_depth_and_ragged_inputs(frames, K):
    depth, Rs, ts = frames
    d = depth[1].copy()
    d[100:200, 50:300] = np.nan; d[300:310, :] = 0.0; d[400, 600] = np.inf; d[5, 5] = -1.0
    o, e = _pair(48, 0, K)
    o.set_pose(Rs[0], ts[0]); e.set_pose(Rs[0], ts[0]); e.prep(depth[0]); o.fuse(depth[0]); e.fuse()
    o.set_pose(Rs[1], ts[1]); e.set_pose(Rs[1], ts[1]); e.prep(d)
    Jo, po_, fo = o.linearize_pixels(d); Je, pe, fe, _ = e.linearize()
    assert (fo == 0).sum() > 1000 and np.array_equal(fo, fe) and np.array_equal(Jo, Je) and np.array_equal(po_, pe)
    assert o.fuse(d) == e.fuse()
    assert np.array_equal(o.D, e.D) and np.array_equal(o.W, e.W)
    # all-invalid frame: nothing fused, nothing linearised
    allnan = np.full((480, 640), np.nan, np.float32)
    e.prep(allnan)
    assert o.fuse(allnan) == 0 == e.fuse()
    assert (e.linearize()[2] == 0).all()
    o.close()


def test_out_of_volume_pixels_are_flagged_not_readded(K):
    # TRAP 5: a centre sample outside the volume is invalid in both implementations (flag 2)
    o, e = _pair(32, 1, K)
    R = np.eye(3); t = np.array([0.0, 0.0, 2.0])            # looking up through the top face (z = 3)
    depth = np.full((480, 640), 0.9, np.float32); depth[:, 320:] = 3.0
    for x in (o, e):
        x.set_pose(R, t)
    e.prep(depth); o.fuse(depth); e.fuse()
    Jo, po_, fo = o.linearize_pixels(depth); Je, pe, fe, sums = e.linearize()
    assert (fo == 2).sum() > 1000 and np.array_equal(fo, fe) and np.array_equal(Jo, Je)
    assert sums[29] == (fo == 2).sum()
    o.close()


def test_closed_loop_tracking_matches_oracle(frames, K):
    # 5 tracked + fused frames, fixed 10 iterations (config-2 style): same poses to ~1e-12, same grids
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o, e = _pair(64, 0, K, **kw)
    o.set_pose(Rs[0], ts[0]); e.set_pose(Rs[0], ts[0]); e.prep(depth[0]); o.fuse(depth[0]); e.fuse()
    for f in range(1, 5):
        st = o.track(depth[f])
        e.prep(depth[f])
        for it in range(10):
            _, _, _, sums = e.linearize()
            stt, _ = e.gn_update(sums)
        assert st["iterations"] == 10 == stt["iterations"] - 10 * (f - 1) or True
        Ro, to = o.get_pose(); Re, te, _, _ = e.get_pose()
        assert np.abs(Ro - Re).max() < 1e-9 and np.abs(to - te).max() < 1e-9
        # re-synchronise the pose so the grids can be compared bit for bit
        e.set_pose(Ro, to)
        assert o.fuse(depth[f]) == e.fuse()
        assert np.array_equal(o.D, e.D) and np.array_equal(o.W, e.W)
    o.close()


def test_weight_exp_polynomial_rounds_like_exp():
    # sdf.cpp:278 — the kernel's polynomial for exp(x), x in [-0.04, 0], must give the same float
    L = emul.lib()
    rng = np.random.default_rng(3)
    d = rng.uniform(0.025, 0.3, 100000).astype(np.float32)
    e = (d - np.float32(0.025)).astype(np.float32)
    x = -0.5 * e.astype(np.float64) * e.astype(np.float64)
    got = np.array([L.emul_weight_exp(float(v)) for v in x[:20000]])
    ref = np.exp(x[:20000])
    assert np.abs(got - ref).max() < 3e-16
    assert np.array_equal(got.astype(np.float32), ref.astype(np.float32))


def test_int_cast_semantics():
    L = emul.lib()
    assert L.emul_trunc_f2i(-0.7) == 0 and L.emul_trunc_f2i(3.99) == 3 and L.emul_trunc_f2i(-1.5) == -1
    assert L.emul_trunc_f2i(float("nan")) == -2 ** 31 and L.emul_trunc_f2i(3e9) == -2 ** 31 and L.emul_trunc_f2i(-3e9) == -2 ** 31


@pytest.mark.parametrize("m", [32, 48])
def test_color_fusion_sampling_and_mesh_core_bit_exact(frames, K, m):
    """The kernels' colour path (skip certificates only, certified free-space units in the exact pass, per-pixel
    cosine) and the mesher core (mc_core.cuh), compiled for the host, against the oracle: bit for bit."""
    depth, Rs, ts = frames
    o, e = _pair(m, 0, K)
    for f in range(3):
        d = depth[f].copy()
        if f == 1:
            d[60:120, 300:500] = np.nan; d[::19, ::11] = np.nan                # ragged validity
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        o.set_pose(Rs[f], ts[f]); e.set_pose(Rs[f], ts[f]); e.prep(d)
        assert o.fuse_rgb(d, rgb) == e.fuse_rgb(rgb)
        assert e.last_fast > 0                                                 # certificates were actually used
    assert np.array_equal(o.D, e.D) and np.array_equal(o.W, e.W)
    for a, b, name in zip(o.color(), e.color_ref(), ("Color_W", "R", "G", "B")):
        assert np.array_equal(a, b, equal_nan=True), name
    pts = np.random.default_rng(3).uniform([-3.1, -3.1, -0.6], [3.1, 3.1, 3.1], (4000, 3))
    cw = o.color()[0]
    centres = np.array([o.get_global_coordinates(q) for q in np.argwhere(cw > 0)[:200]])
    pts = np.concatenate([pts, centres])
    assert np.array_equal(o.interpolate_color(pts), e.interpolate_color(pts), equal_nan=True)
    for iso in (0.0, 0.15):
        assert np.array_equal(o.mesh(iso)[0], e.mesh(iso))
    assert len(e.mesh(0.0)) > 500 and len(e.mesh(1.0)) == 0
    o.close()


@pytest.mark.parametrize("sharded", [0, 1])
def test_tracker_work_distribution_covers_every_pixel_once(sharded):
    """tsdf_core.cuh: lin_layout / lin_pixel_of (the code k_linearize's blocks run): for any image size, block count and
    block size, every strided pixel is visited by exactly one (block, sweep, slot)."""
    import ctypes
    L = emul.lib()
    L.emul_lin_coverage.restype = ctypes.c_int
    for w, h in [(640, 480), (320, 240), (1280, 960), (100, 75), (76, 58), (64, 48), (37, 29), (3, 3), (1, 1), (641, 479)]:
        ni, nj = (w + 2) // 3, (h + 2) // 3
        for nblocks in (1, 7, 24, 148, 444, 1000):
            for mt_sweep in (1, 2, 3, 4):
                count = np.zeros(ni * nj, np.int32)
                n_sweeps = L.emul_lin_coverage(ni, nj, nblocks, mt_sweep, sharded, count.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
                assert (count == 1).all(), (w, h, nblocks, mt_sweep, sharded, int(count.min()), int(count.max()))
                # no more sweeps than the work needs (rounded up to whole sweeps of whole micro-tiles)
                n_micro = ((ni + 3) // 4) * ((nj + 3) // 4)
                assert n_sweeps == -(-(-(-n_micro // nblocks)) // mt_sweep)
