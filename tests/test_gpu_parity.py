"""GPU parity tests: the CUDA path, called through the C ABI (libtsdf_b200.so), against the CPU
oracle on the same seeded synthetic inputs.  Tolerances are BASELINE.json's: voxel D/W 1e-6 abs,
J^T J / J^T r 1e-5 relative (max-norm), poses 1e-4 m / 1e-4 rad.  Where the arithmetic is
fp32/fp64-exact by construction the tests also assert bit equality and say so."""
import ctypes
import os

import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tests.conftest import rot_angle
from tools import synth

pytestmark = pytest.mark.gpu

TOL_DW = 1e-6
TOL_AB = 1e-5
TOL_POSE = 1e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "track_fuse_m32.npz")


def pair(m, metric, K, **kw):
    o = po.Oracle(m=m, use_coord_table=0, metric=metric, **kw)
    o.set_intrinsics(K)
    g = T.Tsdf(T.default_config(m=m, metric=metric, **kw))
    g.set_intrinsics(K)
    return o, g


def ab_rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_library_is_loaded_and_device_present(gpu_lib):
    assert gpu_lib.tsdf_device_count() >= 1
    maps = open("/proc/self/maps").read()
    assert "libtsdf_b200.so" in maps


def test_backproject_and_normals_bit_exact(gpu_lib, frames, K):
    depth, Rs, ts = frames
    o, g = pair(32, 0, K)
    d = depth[2].copy(); d[50:60, 100:200] = np.nan; d[7, 7] = 0.0
    co, no = o.backproject(d); cg, ng = g.backproject(d)
    assert np.array_equal(co, cg, equal_nan=True) and np.array_equal(no, ng, equal_nan=True)
    g.close(); o.close()


@pytest.mark.parametrize("metric,m", [(0, 64), (1, 64), (0, 128)])
def test_fusion_step_parity(gpu_lib, frames, K, metric, m):
    depth, Rs, ts = frames
    o, g = pair(m, metric, K)
    for f in range(4):
        o.set_pose(Rs[f], ts[f])
        n_o = o.fuse(depth[f])
        n_g = g.fuse(depth[f], Rs[f], ts[f])
        assert n_o == n_g
        Dg, Wg = g.download(T.LAYOUT_REFERENCE)
        assert np.abs(Dg - o.D).max() <= TOL_DW and np.abs(Wg - o.W).max() <= TOL_DW
        assert np.array_equal(Dg, o.D) and np.array_equal(Wg, o.W)           # exact by construction
    # the two documented layouts hold the same voxels
    Dx, Wx = g.download(T.LAYOUT_XFASTEST)
    assert np.array_equal(Dx.transpose(2, 1, 0), Dg) and np.array_equal(Wx.transpose(2, 1, 0), Wg)
    g.close(); o.close()


def test_interpolate_distance_parity(gpu_lib, frames, K):
    depth, Rs, ts = frames
    o, g = pair(64, 0, K)
    for f in range(2):
        o.set_pose(Rs[f], ts[f]); o.fuse(depth[f]); g.fuse(depth[f], Rs[f], ts[f])
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-2, 66, (20000, 3)), rng.integers(0, 64, (2000, 3)).astype(float),
                          np.array([[np.nan, 1, 1], [1e12, 3, 3], [-1e12, 3, 3], [63.5, 63.5, 63.5], [-0.5, -0.5, -0.5]])])
    vo, oko = o.interpolate_distance(pts); vg, okg = g.interpolate_distance(pts)
    assert np.array_equal(oko, okg) and np.array_equal(vo, vg, equal_nan=True)
    assert oko.sum() > 1000
    g.close(); o.close()


@pytest.mark.parametrize("metric", [0, 1])
def test_linearisation_parity(gpu_lib, frames, K, metric):
    depth, Rs, ts = frames
    o, g = pair(64, metric, K)
    for f in range(3):
        o.set_pose(Rs[f], ts[f]); o.fuse(depth[f]); g.fuse(depth[f], Rs[f], ts[f])
    o.set_pose(Rs[3], ts[3]); g.set_pose(Rs[3], ts[3])
    Jo, po_, fo = o.linearize_pixels(depth[3]); Jg, pg, fg = g.linearize_pixels(depth[3])
    assert np.array_equal(fo, fg)
    assert np.array_equal(Jo, Jg) and np.array_equal(po_, pg)               # fp32-exact by construction
    A, b, st = o.linearize(depth[3]); Ag, bg, sg = g.linearize(depth[3])
    assert ab_rel(Ag, A) <= TOL_AB and ab_rel(bg, b) <= TOL_AB
    assert ab_rel(Ag, A) <= 1e-12 and ab_rel(bg, b) <= 1e-12               # double sums, only the order differs
    assert sg["n_valid"] == st["n_valid"] and sg["n_oob"] == st["n_oob"] == 0
    assert sg["residual"] == pytest.approx(st["residual"], rel=1e-12)
    # deterministic: the same call twice gives the same bits
    A2, b2, _ = g.linearize(depth[3])
    assert np.array_equal(A2, Ag) and np.array_equal(b2, bg)
    g.close(); o.close()


@pytest.mark.parametrize("wh", [(320, 240), (100, 75), (76, 58), (64, 48), (1280, 960)])
def test_linearisation_parity_other_image_sizes(gpu_lib, wh):
    """The tracker's work distribution (4 x 4 micro-tiles of the strided pixel grid, three per sweep, one block per SM)
    on image sizes whose strided grid is not a multiple of 4, has fewer micro-tiles than blocks, or needs more sweeps
    than the staged slots: per-pixel J / psi / flag bit-equal to the oracle, sums to rounding, one tracked frame."""
    w, h = wh
    s = w / 640.0
    Ks = np.array([525.0 * s, 0.0, (w - 1) / 2.0, 0.0, 525.0 * s, (h - 1) / 2.0, 0.0, 0.0, 1.0])
    depth, Rs, ts = synth.render_sequence(6, K=Ks, w=w, h=h)
    kw = dict(m=64, image_width=w, image_height=h, gauss_newton_max_iteration=6, maximum_twist_diff=float("-inf"))
    o = po.Oracle(use_coord_table=0, **kw); o.set_intrinsics(Ks)
    g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(Ks)
    for f in range(4):
        o.set_pose(Rs[f], ts[f]); o.fuse(depth[f]); g.fuse(depth[f], Rs[f], ts[f])
    o.set_pose(Rs[4], ts[4]); g.set_pose(Rs[4], ts[4])
    Jo, po_, fo = o.linearize_pixels(depth[4]); Jg, pg, fg = g.linearize_pixels(depth[4])
    assert fo.size == ((w + 2) // 3) * ((h + 2) // 3)
    assert np.array_equal(fo, fg) and np.array_equal(Jo, Jg) and np.array_equal(po_, pg)
    A, b, st = o.linearize(depth[4]); Ag, bg, sg = g.linearize(depth[4])
    assert sg["n_valid"] == st["n_valid"] > 0
    assert ab_rel(Ag, A) <= 1e-12 and ab_rel(bg, b) <= 1e-12
    st = o.track(depth[5]); Rg, tg, sg = g.track(depth[5])
    Ro, to = o.get_pose()
    assert sg["iterations"] == st["iterations"] == 6 and sg["n_valid"] == st["n_valid"]
    assert np.abs(tg - to).max() < 1e-9 and np.abs(Rg - Ro).max() < 1e-9
    g.close(); o.close()


@pytest.mark.parametrize("fixed", [True, False])
def test_tracking_parity(gpu_lib, frames, K, fixed):
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf")) if fixed else {}
    o, g = pair(64, 0, K, **kw)
    o.set_pose(Rs[0], ts[0]); o.fuse(depth[0]); g.fuse(depth[0], Rs[0], ts[0])
    for f in range(1, 5):
        st = o.track(depth[f])
        Rg, tg, sg = g.track(depth[f])
        Ro, to = o.get_pose()
        assert sg["iterations"] == st["iterations"] and sg["stopped"] == st["stopped"] and sg["n_valid"] == st["n_valid"]
        assert np.linalg.norm(tg - to) <= TOL_POSE and rot_angle(Rg, Ro) <= TOL_POSE
        assert np.abs(tg - to).max() < 1e-9 and np.abs(Rg - Ro).max() < 1e-9
        R2, t2 = g.get_pose()
        assert np.array_equal(R2, Rg) and np.array_equal(t2, tg)
        # resynchronise so the comparison stays step-wise; grids stay bit-identical
        o.fuse(depth[f]); g.fuse(depth[f], Ro, to)
    Dg, Wg = g.download()
    assert np.array_equal(Dg, o.D) and np.array_equal(Wg, o.W)
    g.close(); o.close()


def test_closed_loop_free_running(gpu_lib, frames, K):
    """track_and_fuse with no resynchronisation, 8 frames at 128^3.  Reduction order makes the
    poses differ at ~1e-12, which can flip a handful of (int) pixel truncations in fusion
    (DESIGN.md 'closed-loop parity'), so the grid test is statistical; the poses meet 1e-4."""
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o, g = pair(128, 0, K, **kw)
    o.set_pose(Rs[0], ts[0]); o.fuse(depth[0]); g.fuse(depth[0], Rs[0], ts[0])
    for f in range(1, 9):
        o.track(depth[f]); n_o = o.fuse(depth[f])
        Rg, tg, sg, n_g = g.track_and_fuse(depth[f])
        Ro, to = o.get_pose()
        assert np.linalg.norm(tg - to) <= TOL_POSE and rot_angle(Rg, Ro) <= TOL_POSE
        assert abs(n_g - n_o) <= 1e-4 * n_o
        assert np.linalg.norm(tg - ts[f]) < 0.06
    Dg, Wg = g.download()
    bad = (np.abs(Dg - o.D) > TOL_DW) | (np.abs(Wg - o.W) > TOL_DW)
    assert bad.mean() < 1e-5
    g.close(); o.close()


def test_async_stream_path_equals_sync_path(gpu_lib, frames, K):
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    a = T.Tsdf(T.default_config(m=64, **kw)); b = T.Tsdf(T.default_config(m=64, **kw))
    for x in (a, b):
        x.set_intrinsics(K); x.set_pose(Rs[0], ts[0])
    n = 6
    dev = a.dev_alloc(depth[:n].nbytes)
    a.dev_upload(dev, depth[:n])
    fb = depth[0].nbytes
    a.enqueue_frame(dev, track=0, slot=0)
    for f in range(1, n):
        a.enqueue_frame(dev + f * fb, track=1, slot=f)
    a.sync()
    b.fuse(depth[0])
    for f in range(1, n):
        Rb, tb, sb, _ = b.track_and_fuse(depth[f])
        Ra, ta, sa = a.read_pose_ring(f)
        assert np.array_equal(Ra, Rb) and np.array_equal(ta, tb) and sa["iterations"] == 10
    Da, Wa = a.download(); Db, Wb = b.download()
    assert np.array_equal(Da, Db) and np.array_equal(Wa, Wb)
    a.dev_free(dev)
    assert a.kernel_launch_count() >= 1 + n + 10 * (n - 1) + n
    a.close(); b.close()


def test_submit_frame_buffer_ring_may_be_recycled_as_documented(gpu_lib, frames, K):
    """tsdf_submit_frame: "the host buffer must stay valid until 4 later submissions returned".  A caller that takes
    the contract literally keeps a ring of FIVE pinned buffers and overwrites buffer n % 5 right after submission
    n + 4 has returned — far ahead of the GPU unless the library makes the host wait.  Poses and grid must equal the
    run with one distinct buffer per frame (the lifetime rule was documented but not enforced)."""
    from tracking_sdf_b200 import capi
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    a = T.Tsdf(T.default_config(m=128, **kw)); b = T.Tsdf(T.default_config(m=128, **kw))
    for x in (a, b):
        x.set_intrinsics(K); x.set_pose(Rs[0], ts[0])
    n, NR = 12, 5
    ring = capi.pinned_empty((NR,) + depth[0].shape, np.float32)
    for f in range(n):
        ring[f % NR][...] = depth[f]                      # overwrite the slot whose 4-later submission has returned
        a.submit_frame(ring[f % NR], track=0 if f == 0 else 1, slot=f)
        ring[f % NR - 4][...] = np.nan if f >= 4 else ring[f % NR - 4]   # poison the buffer of submission f - 4: it must be consumed by now
    a.sync()
    b.fuse(depth[0])
    for f in range(1, n):
        Rb, tb, sb, _ = b.track_and_fuse(depth[f])
        Ra, ta, sa = a.read_pose_ring(f)
        assert np.array_equal(Ra, Rb) and np.array_equal(ta, tb), f
    Da, Wa = a.download(); Db, Wb = b.download()
    assert np.array_equal(Da, Db) and np.array_equal(Wa, Wb)
    a.close(); b.close()


def test_host_streaming_path_equals_sync_path(gpu_lib, frames, K):
    """tsdf_submit_frame (copy stream + device frame ring, H2D of frame n+1 under compute of frame n)
    gives the same poses and grid as the synchronous host-buffer call."""
    from tracking_sdf_b200 import capi
    depth, Rs, ts = frames
    kw = dict(gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    a = T.Tsdf(T.default_config(m=64, **kw)); b = T.Tsdf(T.default_config(m=64, **kw))
    for x in (a, b):
        x.set_intrinsics(K); x.set_pose(Rs[0], ts[0])
    n = 10                                           # more frames than the 4-deep stage ring
    pinned = capi.pinned_empty((n, 480, 640), np.float32)
    pinned[:] = depth[:n]
    a.submit_frame(pinned[0], track=0, slot=0)
    for f in range(1, n):
        a.submit_frame(pinned[f], track=1, slot=f)
    a.sync()
    b.fuse(depth[0])
    for f in range(1, n):
        Rb, tb, sb, _ = b.track_and_fuse(depth[f])
        Ra, ta, sa = a.read_pose_ring(f)
        assert np.array_equal(Ra, Rb) and np.array_equal(ta, tb)
    Da, Wa = a.download(); Db, Wb = b.download()
    assert np.array_equal(Da, Db) and np.array_equal(Wa, Wb)
    a.close(); b.close()


def test_against_committed_golden(gpu_lib, frames):
    gold = np.load(GOLD)
    depth, Rs, ts = frames
    for metric in (0, 1):
        gg = lambda k: gold["m%d_%s" % (metric, k)]
        g = T.Tsdf(T.default_config(m=32, metric=metric, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf")))
        g.set_intrinsics(gold["K"])
        for f in range(3):
            assert g.fuse(depth[f], Rs[f], ts[f]) == gg("n_updated")[f]
        D, W = g.download()
        assert np.abs(D - gg("D")).max() <= TOL_DW and np.abs(W - gg("W")).max() <= TOL_DW
        g.set_pose(Rs[3], ts[3])
        A, b, st = g.linearize(depth[3])
        assert ab_rel(A, gg("A")) <= TOL_AB and ab_rel(b, gg("b")) <= TOL_AB
        J, psi, flag = g.linearize_pixels(depth[3])
        assert np.array_equal(flag, gg("flag")) and np.array_equal(J[::16], gg("J")) and np.array_equal(psi[::16], gg("psi"))
        g.set_pose(Rs[2], ts[2])
        R, t, _ = g.track(depth[3])
        assert np.linalg.norm(t - gg("t_tracked")) <= TOL_POSE and rot_angle(R, gg("R_tracked")) <= TOL_POSE
        v, ok = g.interpolate_distance(gg("sample_pts"))
        assert np.array_equal(ok, gg("sample_ok")) and np.array_equal(v, gg("sample_val"), equal_nan=True)
        g.close()
    g = T.Tsdf(T.default_config(m=32))
    for tw, R, t in zip(gold["exp_twist"], gold["exp_R"], gold["exp_t"]):
        Rg, tg = g.exp_map(tw)
        assert np.abs(Rg - R).max() < 1e-14 and np.abs(tg - t).max() < 1e-14
    g.close()


def test_edge_cases(gpu_lib, frames, K):
    depth, Rs, ts = frames
    # no intrinsics -> status, not exit(0) (sdf.cpp:227-229)
    g = T.Tsdf(T.default_config(m=32))
    with pytest.raises(T.TsdfError) as e:
        g.fuse(depth[0])
    assert e.value.status == 2
    g.set_intrinsics(K)
    # initial pose and grid (camera_tracking.cpp:5-8, sdf.cpp:29-31)
    R, t = g.get_pose()
    assert np.array_equal(R, [[1, 0, 0], [0, 0, -1], [0, -1, 0]]) and np.array_equal(t, [0, 0, 1])
    D, W = g.download()
    assert (D == np.float32(15.5)).all() and (W == 0).all()
    # empty volume: singular normal equations are reported, pose kept (TRAP 12)
    with pytest.raises(T.TsdfError) as e:
        g.track(depth[0])
    assert e.value.status == 4
    R1, t1 = g.get_pose()
    assert np.array_equal(R, R1) and np.array_equal(t, t1)
    # all-invalid frame fuses nothing
    assert g.fuse(np.full((480, 640), np.nan, np.float32), Rs[0], ts[0]) == 0
    g.close()
    # ragged validity + skewed K, both metrics of the data path
    d = depth[1].copy(); d[100:200, 50:300] = np.nan; d[300:310, :] = 0.0; d[400, 600] = np.inf
    Ks = K.copy(); Ks[1] = 0.7
    o, g = pair(48, 0, Ks)
    o.set_pose(Rs[0], ts[0]); o.fuse(depth[0]); g.fuse(depth[0], Rs[0], ts[0])
    o.set_pose(Rs[1], ts[1]); g.set_pose(Rs[1], ts[1])
    Jo, po_, fo = o.linearize_pixels(d); Jg, pg, fg = g.linearize_pixels(d)
    assert np.array_equal(fo, fg) and np.array_equal(Jo, Jg) and (fo == 0).sum() > 1000
    assert o.fuse(d) == g.fuse(d)
    Dg, Wg = g.download()
    assert np.array_equal(Dg, o.D) and np.array_equal(Wg, o.W)
    # out-of-volume centres are flagged (TRAP 5)
    o2, g2 = pair(32, 1, K)
    dd = np.full((480, 640), 0.9, np.float32); dd[:, 320:] = 3.0
    Rz, tz = np.eye(3), np.array([0.0, 0.0, 2.0])
    o2.set_pose(Rz, tz); o2.fuse(dd); g2.fuse(dd, Rz, tz)
    _, _, f2o = o2.linearize_pixels(dd); _, _, f2g = g2.linearize_pixels(dd)
    assert np.array_equal(f2o, f2g) and (f2g == 2).sum() > 1000
    A, b, st = g2.linearize(dd)
    assert st["n_oob"] == (f2g == 2).sum()
    for x in (o, g, o2, g2):
        x.close()


def test_upload_download_round_trip(gpu_lib, K):
    g = T.Tsdf(T.default_config(m=32))
    rng = np.random.default_rng(9)
    D = rng.normal(size=(32, 32, 32)).astype(np.float32); W = rng.uniform(0, 3, (32, 32, 32)).astype(np.float32)
    g.upload(D, W, T.LAYOUT_REFERENCE)
    D2, W2 = g.download(T.LAYOUT_REFERENCE)
    assert np.array_equal(D, D2) and np.array_equal(W, W2)
    Dx, Wx = g.download(T.LAYOUT_XFASTEST)
    assert np.array_equal(Dx, D.transpose(2, 1, 0)) and np.array_equal(Wx, W.transpose(2, 1, 0))
    g.upload(Dx, Wx, T.LAYOUT_XFASTEST)
    D3, _ = g.download(T.LAYOUT_REFERENCE)
    assert np.array_equal(D3, D)
    # accessor maps (sdf.h:113-157)
    assert g.get_array_index(1, 2, 3) == 32 * 32 + 64 + 3 and g.get_array_index(32, 0, 0) == -1
    assert tuple(g.get_voxel_coordinates_idx(32 * 32 + 64 + 3)) == (1, 2, 3)
    v, ok = g.interpolate_distance([[4.0, 5.0, 6.0]])
    assert ok[0] == (W[4, 5, 6] > 0) and (not ok[0] or v[0] == D[4, 5, 6])
    g.close()


def test_reference_class_mirror(gpu_lib, frames, K):
    """The SDF / CameraTracking mirror (tracking_sdf_b200/api.py) used the way the ROS node uses
    the reference classes (sdf_reconstruction.cpp:82-91, 66-74)."""
    depth, Rs, ts = frames
    sdf = T.SDF(64, 6.0, 6.0, 3.5, (-3.0, -3.0, -0.5), 0.3, 0.025)
    cam = T.CameraTracking(20, 0.001, 1.0, 0.01, sdf)
    with pytest.raises(T.TsdfError):
        sdf.update(cam, depth[0])                      # K not received yet
    cam.camera_info_cb(K.reshape(3, 3))
    cam.set_camera_transformation(Rs[0], ts[0])
    o = po.Oracle(m=64, use_coord_table=0); o.set_intrinsics(K); o.set_pose(Rs[0], ts[0])
    assert sdf.update(cam, depth[0]) == o.fuse(depth[0])
    cam.estimate_new_position(sdf, depth[1]); o.track(depth[1])
    assert np.abs(cam.trans - o.get_pose()[1]).max() < 1e-9 and np.abs(cam.rot - o.get_pose()[0]).max() < 1e-9
    assert np.allclose(cam.rot_inv @ cam.rot, np.eye(3), atol=1e-12)
    assert sdf.get_number_of_voxels() == 64 ** 3 and sdf.get_array_index((1, 2, 3)) == 64 * 64 + 128 + 3
    # the colour half of update and the visualisation thread's calls (sdf.cpp:294-304, 327-385)
    o.set_pose(cam.rot, cam.trans)
    rgb = synth.synth_rgb(depth[1], Rs[1], ts[1])
    assert sdf.update(cam, depth[1], rgb) == o.fuse_rgb(depth[1], rgb)
    pts, rgba = sdf.mesh(colors=True)
    xo, wo, co = o.mesh(0.0, world=True, colors=True)
    assert np.array_equal(pts, wo) and np.array_equal(rgba, co, equal_nan=True)
    assert np.array_equal(sdf.interpolate_color(wo[:50]), co[:50], equal_nan=True)
    assert np.array_equal(sdf.Color[1], o.color()[1])
    o.close()


def test_fast_reciprocal_is_exact_on_its_range(gpu_lib):
    """The tracker computes w = 1/volume (sdf.cpp:154) with MUFU.RCP + one FMA Newton step.  Every
    float in [2^-17, 4] (volume lies in (1e-5, 3]) is checked against IEEE 1.0f/x on the device."""
    g = T.Tsdf(T.default_config(m=32))
    assert g.debug_check_rcp(2.0 ** -17, 4.0) == 0
    g.close()


def test_fusion_weight_exp_exhaustive(gpu_lib):
    """sdf.cpp:278: w = (float)exp(-0.5 (d - eps)^2) for eps <= d <= delta, i.e. e = d - eps in [0, 0.275].  The device
    evaluates the exponential with a degree-9 polynomial in double (x >= -0.04) or CUDA's exp.  EVERY float e in
    [0, 0.3] (1.05e9 operands) is checked on the device: the weight is provably the correctly rounded float of the
    true exponential except for the listed operands (true value within 6e-16 relative of a float rounding
    boundary); those are compared here with the host libm the reference calls.  What remains is the exact number
    of operands, out of 1.05e9, where device and reference can differ — by one float ulp (6e-8) of a weight near 1."""
    import math
    g = T.Tsdf(T.default_config(m=32))
    n, e, w = g.debug_check_weight_exp(0.0, 0.3)
    g.close()
    assert n == len(e) and n < 200                               # ~1e9 * 2 * 6e-16 / 2^-24 expected: a few dozen
    w_host = np.array([np.float32(math.exp(-0.5 * float(x) * float(x))) for x in e], np.float32)
    differ = int((w_host != w).sum())
    ulp = np.abs(w_host.view(np.int32) - w.view(np.int32)).max() if n else 0
    print("weight_exp: %d ambiguous of ~1.05e9 operands, %d differ from the host libm, max %d float ulp" % (n, differ, ulp))
    assert differ <= 8 and ulp <= 1                              # <= 6e-8 relative in w: far inside the 1e-6 on D/W


def test_certificates_are_certified(gpu_lib, frames, K):
    """Fusion decides whole four-voxel units from the certificate pyramid (free space: updated with
    d = -delta, w = 1; or skipped).  The self-check build runs the exact fp64 path on EVERY voxel of
    every certified unit and compares: trajectory frames at 256^3 and the dense pose, both metrics."""
    depth, Rs, ts = frames
    for metric in (0, 1):
        g = T.Tsdf(T.default_config(m=256, metric=metric)); g.set_intrinsics(K)
        for f in (0, 3, 7, 11):
            g.set_pose(Rs[f], ts[f])
            r = g.debug_fuse_check(depth[f])
            assert r["wrong"] == 0 and r["fast"] > 100000, r
        d = depth[5].copy(); d[100:300, 200:500] = np.nan; d[::7, ::5] = 0.0      # ragged validity
        g.set_pose(Rs[5], ts[5])
        r = g.debug_fuse_check(d)
        assert r["wrong"] == 0, r
        R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float)
        g.set_pose(R, [0.0, -12.0, 1.25])
        r = g.debug_fuse_check(np.full((480, 640), 40.0, np.float32))
        assert r["wrong"] == 0 and r["fast"] >= 0.99 * 256 ** 3, r
        D, W = g.download()
        assert (W == 0).all()                      # the self-check never writes
        g.close()


def test_64bit_index_path_of_the_tracker(gpu_lib, frames, K, monkeypatch):
    """Stores of 2^32 voxels and more (>= 32 GiB) take the tracker's 64-bit index path; force it on a small
    store and compare with the 32-bit path: identical normal equations and pose."""
    depth, Rs, ts = frames
    kw = dict(m=64, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    out = []
    for force in ("0", "1"):
        monkeypatch.setenv("TSDF_B200_IDX64", force)
        g = T.Tsdf(T.default_config(**kw)); g.set_intrinsics(K)
        g.fuse(depth[0], Rs[0], ts[0])
        A, b, st = g.linearize(depth[1])
        R, t, st2 = g.track(depth[1])
        out.append((A, b, R, t))
        g.close()
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(x, y)
