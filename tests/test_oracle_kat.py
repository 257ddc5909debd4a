"""Known-answer tests that pin the ORACLE itself.

The reference has no tests or golden vectors and cannot be compiled here, so the restatement
is pinned from first principles: each test states a fact that follows directly from the
reference source lines cited.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tools import synth


@pytest.fixture(scope="module")
def small():
    o = po.Oracle(m=32, use_coord_table=0)
    o.set_intrinsics(synth.K_DEFAULT)
    yield o
    o.close()


def test_fp32_constants_match_survey_table():
    # SURVEY.md §8 "derived float constants" (sdf.cpp:19-21, camera_tracking.cpp:13-17)
    o = po.Oracle(m=256, use_coord_table=0)
    c = o.constants()
    assert c[0] == np.float32(42.66666793823242) and c[2] == np.float32(73.14286041259766)
    assert c[3] == np.float32(0.046875) and c[5] == np.float32(0.02734374813735485)
    o.close()


def test_index_round_trip(small):
    # sdf.h:113-136: get_array_index o get_voxel_coordinates(int) = id; z-fastest; -1 outside
    m = small.m
    for idx in [0, 1, m - 1, m, m * m, m * m * m - 1, 12345]:
        i, j, k = small.get_voxel_coordinates_idx(idx)
        assert small.get_array_index(i, j, k) == idx
    assert small.get_array_index(1, 2, 3) == m * m * 1 + m * 2 + 3
    for bad in [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (m, 0, 0), (0, m, 0), (0, 0, m)]:
        assert small.get_array_index(*bad) == -1


def test_coordinate_round_trip(small):
    # sdf.h:143-157: voxel centres sit at integer voxel coordinates
    rng = np.random.default_rng(0)
    for _ in range(200):
        ijk = rng.integers(0, small.m, 3)
        v = small.get_voxel_coordinates(small.get_global_coordinates(ijk))
        assert np.abs(v - ijk).max() < 1e-5       # fp32 scale factors: not exactly integer
    g = small.get_global_coordinates([0, 0, 0])
    np.testing.assert_allclose(g, [-3 + 0.5 * 6 / 32, -3 + 0.5 * 6 / 32, -0.5 + 0.5 * 3.5 / 32], rtol=1e-7)


def test_initial_state(small):
    # sdf.cpp:29-31 and camera_tracking.cpp:5-8
    small.reset()
    assert (small.D == np.float32(15.5)).all() and (small.W == 0).all()
    o = po.Oracle(m=8, use_coord_table=0)
    R, t = o.get_pose()
    np.testing.assert_array_equal(R, [[1, 0, 0], [0, 0, -1], [0, -1, 0]])
    np.testing.assert_array_equal(t, [0, 0, 1])
    o.close()


def test_interpolate_exact_centre_returns_D(small):
    # sdf.cpp:151-153: volume < 1e-5 -> return D[idx]
    small.reset()
    small.D[5, 6, 7] = 0.123
    small.W[5, 6, 7] = 1.0
    v, ok = small.interpolate_distance([[5.0, 6.0, 7.0]])
    assert ok[0] and v[0] == np.float32(0.123)


def test_interpolate_all_unseen_is_not_interpolated(small):
    # sdf.cpp:139,149,162: no neighbour with W>0 -> is_interpolated false, 0/0 = NaN
    small.reset()
    v, ok = small.interpolate_distance([[5.3, 6.2, 7.9]])
    assert not ok[0] and np.isnan(v[0])


def test_interpolate_cell_centre_is_plain_mean(small):
    # at (i+.5, j+.5, k+.5) all eight L1 distances are 1.5 -> equal weights
    small.reset()
    rng = np.random.default_rng(1)
    vals = rng.uniform(-0.3, 0.3, (2, 2, 2)).astype(np.float32)
    small.D[10:12, 10:12, 10:12] = vals
    small.W[10:12, 10:12, 10:12] = 1.0
    v, ok = small.interpolate_distance([[10.5, 10.5, 10.5]])
    assert ok[0]
    np.testing.assert_allclose(v[0], vals.astype(np.float64).mean(), rtol=2e-6)


def test_interpolate_is_inverse_l1_not_trilinear(small):
    # TRAP 1 (sdf.cpp:146,154): weights 1/L1; hand-computed for a point off-centre
    small.reset()
    small.D[3:5, 3:5, 3:5] = 0.0
    small.W[3:5, 3:5, 3:5] = 1.0
    small.D[3, 3, 3] = 1.0
    p = np.array([3.25, 3.5, 3.75])
    f = p.astype(np.float32)
    w_sum = np.float32(0); s = np.float32(0)
    for io in (0, 1):
        for jo in (0, 1):
            for ko in (0, 1):
                vol = abs(np.float32(3 + io) - f[0]) + abs(np.float32(3 + jo) - f[1]) + abs(np.float32(3 + ko) - f[2])
                w = np.float32(1.0 / np.float64(vol))
                w_sum = np.float32(w_sum + w)
                s = np.float32(s + np.float32(w * (1.0 if (io, jo, ko) == (0, 0, 0) else 0.0)))
    v, ok = small.interpolate_distance([p])
    assert ok[0] and v[0] == np.float32(s / w_sum)
    trilinear = 0.75 * 0.5 * 0.25
    assert abs(v[0] - trilinear) > 1e-2


def test_interpolate_partial_neighbourhood_renormalises(small):
    # TRAP 11: is_interpolated if ANY neighbour has W>0; weights renormalise over the valid subset
    small.reset()
    small.D[8, 8, 8] = 0.2
    small.W[8, 8, 8] = 3.0
    v, ok = small.interpolate_distance([[8.4, 8.3, 8.2]])
    assert ok[0] and v[0] == pytest.approx(0.2, rel=1e-6)


def test_interpolate_truncates_toward_zero(small):
    # TRAP 9 (sdf.cpp:143-145): (int) truncation -> coordinates in (-1,0) use cells {0,1}
    small.reset()
    small.D[0, 0, 0] = 0.1
    small.W[0, 0, 0] = 1.0
    v, ok = small.interpolate_distance([[-0.5, -0.5, -0.5]])
    assert ok[0] and v[0] == np.float32(0.1)


def test_float_division_equals_double_division_rounded():
    # the CUDA core computes w = 1.0f/volume; the reference (float)(1.0/volume) (sdf.cpp:154).
    # Equal for all floats because 53 >= 2*24+2 (double rounding is innocuous for division).
    rng = np.random.default_rng(2)
    vol = rng.uniform(1e-5, 3.0, 200000).astype(np.float32)
    a = (np.float32(1.0) / vol)
    b = (1.0 / vol.astype(np.float64)).astype(np.float32)
    assert np.array_equal(a, b)


def test_vol_exact_threshold_constant():
    # `volume < 0.00001` with volume float promoted to double (sdf.cpp:151) <=> volume <= float(1e-5)
    c = np.float32(1e-5)
    assert float(c) < 1e-5 and float(np.nextafter(c, np.float32(1))) >= 1e-5


def test_exp_map_known_values():
    # eigen_utils.cpp:40-128
    R, t = po.exp_map(np.zeros(6))
    np.testing.assert_array_equal(R, np.eye(3)); np.testing.assert_array_equal(t, np.zeros(3))
    R, t = po.exp_map([0.1, -0.2, 0.3, 0, 0, 0])            # pure translation: dt = v
    np.testing.assert_array_equal(R, np.eye(3)); np.testing.assert_allclose(t, [0.1, -0.2, 0.3], rtol=1e-15)
    R, t = po.exp_map([0, 0, 0, 0, 0, np.pi / 2])            # quarter turn about z
    np.testing.assert_allclose(R, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)
    # small-angle guards: theta < 2.5e-4 -> mcosc = 1/2, msinc = 1/6; theta < 1e-8 -> sinc = 1
    R, t = po.exp_map([1, 0, 0, 0, 0, 1e-9])
    assert R[0, 0] == pytest.approx(1.0, abs=1e-15) and t[0] == pytest.approx(1.0, rel=1e-15)
    R, t = po.exp_map([1, 2, 3, 1e-4, 0, 0])
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-8
    # agreement with the closed form for a generic twist (twist order (v, w), TRAP 7)
    w = np.array([0.3, -0.2, 0.5]); v = np.array([0.1, 0.4, -0.7])
    th = np.linalg.norm(w); Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    Rref = np.eye(3) + np.sin(th) / th * Wx + (1 - np.cos(th)) / th ** 2 * Wx @ Wx
    Vref = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Wx + (th - np.sin(th)) / th ** 3 * Wx @ Wx
    R, t = po.exp_map(np.concatenate([v, w]))
    np.testing.assert_allclose(R, Rref, atol=1e-14); np.testing.assert_allclose(t, Vref @ v, atol=1e-14)


def test_pose_inverse(small):
    # camera_tracking.cpp:59-65
    _, R, t = synth.load_trajectory()
    small.set_pose(R[10], t[10])
    Ri, ti = small.get_pose_inv()
    np.testing.assert_allclose(Ri @ R[10], np.eye(3), atol=1e-12)
    np.testing.assert_allclose(ti, -Ri @ t[10], atol=1e-15)


def _single_voxel_case(metric, plane_z, m=32):
    """Camera at the world origin looking along +z (R = I): fuse a fronto-parallel plane."""
    o = po.Oracle(m=m, use_coord_table=0, metric=metric, origin=(-3.0, -3.0, -0.5))
    o.set_intrinsics(synth.K_DEFAULT)
    o.set_pose(np.eye(3), np.zeros(3))
    depth = np.full((480, 640), plane_z, np.float32)
    return o, depth


@pytest.mark.parametrize("metric", [0, 1])
def test_fusion_by_hand(metric):
    # sdf.cpp:245-292 for voxels on the optical axis of an axis-aligned camera.
    o, depth = _single_voxel_case(metric, 2.0)
    n1 = o.fuse(depth)
    assert n1 > 0
    m = o.m
    i = j = m // 2                      # voxel centre x = y = +0.09375 m: projects near the principal point
    delta, eps = np.float32(0.3), np.float32(0.025)
    seen = {"front": 0, "band": 0, "plateau": 0, "behind": 0}
    for k in range(m):
        g = o.get_global_coordinates([i, j, k])
        if g[2] < 0:
            assert o.W[i, j, k] == 0          # behind the camera, sdf.cpp:247
            continue
        uu = 525.0 * g[0] / g[2] + 319.5; vv = 525.0 * g[1] / g[2] + 239.5
        if not (1 <= uu < 639 and 1 <= vv < 479):
            assert o.W[i, j, k] == 0          # projects outside the image (or onto the normal-less border), sdf.cpp:254,260
            continue
        d = np.float32(g[2] - 2.0)            # both metrics agree for a fronto-parallel plane (n = (0,0,-1))
        if d > delta:
            assert o.W[i, j, k] == 0 and o.D[i, j, k] == np.float32(15.5); seen["behind"] += 1   # sdf.cpp:280-283
        elif d < -delta:
            assert o.W[i, j, k] == 1 and o.D[i, j, k] == -delta; seen["front"] += 1              # sdf.cpp:285-287
        elif d >= eps:
            w = np.float32(np.exp(-0.5 * np.float64(np.float32(d - eps)) * np.float64(np.float32(d - eps))))
            assert o.W[i, j, k] == w; seen["band"] += 1                                           # sdf.cpp:276-279
            assert o.D[i, j, k] == np.float32(np.float32(w * d) / w)
        else:
            assert o.W[i, j, k] == 1 and o.D[i, j, k] == pytest.approx(d, abs=1e-7); seen["plateau"] += 1
    assert all(v > 0 for v in seen.values()), seen
    # second observation of a nearer plane: running weighted mean, sdf.cpp:289-292
    D1, W1 = o.D.copy(), o.W.copy()
    o.fuse(np.full((480, 640), 1.9, np.float32))
    k = int(np.argmin([abs(o.get_global_coordinates([i, j, kk])[2] - 1.95) for kk in range(m)]))
    g = o.get_global_coordinates([i, j, k])
    d2 = np.float32(g[2] - np.float64(np.float32(1.9)))
    w2 = np.float32(1.0) if d2 < eps else np.float32(np.exp(-0.5 * np.float64(np.float32(d2 - eps)) ** 2))
    Wn = np.float32(W1[i, j, k] + w2)
    Dn = np.float32(np.float32(np.float32(W1[i, j, k] * D1[i, j, k]) + np.float32(w2 * d2)) / Wn)
    assert o.W[i, j, k] == Wn and o.D[i, j, k] == pytest.approx(Dn, abs=1e-7)
    o.close()


def test_fusion_pixel_truncation_and_bounds():
    # sdf.cpp:251-254: (int) truncation; u in (-1,0) lands in column 0; outside image skipped
    o, depth = _single_voxel_case(1, 2.5, m=64)
    o.fuse(depth)
    W = o.W
    upd = np.argwhere(W > 0)
    K = synth.K_DEFAULT
    g = np.array([o.get_global_coordinates(v) for v in upd[::37]])
    u = K[0] * g[:, 0] / g[:, 2] + K[2]; v = K[4] * g[:, 1] / g[:, 2] + K[5]
    assert (u > -1).all() and (u < 640).all() and (v > -1).all() and (v < 480).all() and (g[:, 2] >= 0).all()
    o.close()


def test_fusion_skips_nan_depth_and_normals():
    # sdf.cpp:260: NaN point or normal -> voxel untouched. Depth discontinuity -> NaN normal (plane metric only)
    o, depth = _single_voxel_case(0, 2.0)
    depth[:, :320] = np.nan
    o.fuse(depth)
    o2, depth2 = _single_voxel_case(0, 2.0)
    o2.fuse(depth2)
    assert 0 < (o.W > 0).sum() < (o2.W > 0).sum()
    cloud, normals = o.backproject(depth)
    assert np.isnan(cloud[:, :320]).all() and np.isnan(normals[:, :321]).all()    # column 320 lost its left neighbour
    assert not np.isnan(normals[1:-1, 322:-1]).any() and np.isnan(normals[0]).all()
    np.testing.assert_allclose(normals[100, 400], [0, 0, -1], atol=1e-6)          # toward the camera
    o.close(); o2.close()


def test_sphere_gradient_and_gauss_newton_pull():
    # sphere-filled volume (sdf.cpp:99-126). J_trans ~ unit radial direction;
    # GN from a perturbed pose moves toward the truth.
    m = 64
    o = po.Oracle(m=m, use_coord_table=0, gauss_newton_max_iteration=20, maximum_twist_diff=float("-inf"))
    o.set_intrinsics(synth.K_DEFAULT)
    c = np.array([0.0, 0.0, 1.25]); r = 0.8
    o.create_circle(r, *c)
    # camera 2.5 m from the sphere centre looking at it along world +y
    R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], float); t = np.array([0.0, -2.5, 1.25])
    # synthetic depth of that sphere from the true pose
    u, v = np.meshgrid(np.arange(640), np.arange(480))
    dc = np.stack([(u - 319.5) / 525, (v - 239.5) / 525, np.ones_like(u, float)], -1)
    dw = dc @ R.T
    oc = t - c
    a = (dw * dw).sum(-1); b = dw @ oc; cc = oc @ oc - r * r
    disc = b * b - a * cc
    lam = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / a, np.nan)
    depth = lam.astype(np.float32)
    o.set_pose(R, t)
    J, psi, flag = o.linearize_pixels(depth)
    ok = flag == 1
    assert ok.sum() > 1000
    assert np.abs(psi[ok]).max() < 0.05                         # on-surface points: SDF ~ 0 (voxel 9 cm)
    # radial direction at each valid pixel
    s = 3
    ii, jj = np.meshgrid(np.arange(0, 640, s), np.arange(0, 480, s), indexing="ij")
    lamv = lam[jj, ii].reshape(-1)[ok]
    pw = (dc[jj, ii].reshape(-1, 3)[ok] * lamv[:, None]) @ R.T + t
    rad = (pw - c) / np.linalg.norm(pw - c, axis=1, keepdims=True)
    cosang = (J[ok, :3] * rad).sum(1) / np.linalg.norm(J[ok, :3], axis=1)
    assert np.median(cosang) > 0.97
    # perturb and track
    t_bad = t + np.array([0.03, -0.02, 0.025])
    o.set_pose(R, t_bad)
    st = o.track(depth)
    _, t_new = o.get_pose()
    assert st["iterations"] == 20 and not st["singular"]
    assert np.linalg.norm(t_new - t) < 0.5 * np.linalg.norm(t_bad - t)
    o.close()


def test_signed_stop_test_and_fixed_iterations(frames, K):
    # TRAP 8 (camera_tracking.cpp:216-224): signed comparison, the update of the stopping iteration is applied
    depth, Rs, ts = frames
    o = po.Oracle(m=64, use_coord_table=0)
    o.set_intrinsics(K); o.set_pose(Rs[0], ts[0]); o.fuse(depth[0])
    st = o.track(depth[1])
    assert 1 <= st["iterations"] <= 20
    o2 = po.Oracle(m=64, use_coord_table=0, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o2.set_intrinsics(K); o2.set_pose(Rs[0], ts[0]); o2.fuse(depth[0])
    st2 = o2.track(depth[1])
    assert st2["iterations"] == 10 and st2["stopped"] == 0 and st2["n_oob"] == 0
    o.close(); o2.close()


def test_singular_system_is_reported_not_nan(K):
    # TRAP 12: empty volume -> no pixel interpolates -> A = 0; the reference would produce a NaN pose
    o = po.Oracle(m=32, use_coord_table=0)
    o.set_intrinsics(K)
    R0, t0 = o.get_pose()
    st = o.track(np.full((480, 640), 1.0, np.float32))
    R1, t1 = o.get_pose()
    assert st["singular"] == 1 and st["n_valid"] == 0
    np.testing.assert_array_equal(R0, R1); np.testing.assert_array_equal(t0, t1)
    o.close()


def test_coord_table_and_on_the_fly_agree(frames, K):
    # sdf.cpp:40-41,244: the precomputed global_coords table holds exactly get_global_coordinates
    depth, Rs, ts = frames
    a = po.Oracle(m=32, use_coord_table=1); b = po.Oracle(m=32, use_coord_table=0)
    for o in (a, b):
        o.set_intrinsics(K); o.set_pose(Rs[0], ts[0]); o.fuse(depth[0])
    assert np.array_equal(a.D, b.D) and np.array_equal(a.W, b.W)
    a.close(); b.close()


def test_tracking_follows_ground_truth(frames, K):
    # closed loop on the synthetic sequence stays near the GT path (cm level at 128^3)
    depth, Rs, ts = frames
    o = po.Oracle(m=128, use_coord_table=0)
    o.set_intrinsics(K); o.set_pose(Rs[0], ts[0]); o.fuse(depth[0])
    for f in range(1, 6):
        st = o.track(depth[f])
        assert st["n_oob"] == 0 and not st["singular"]
        o.fuse(depth[f])
        _, t = o.get_pose()
        assert np.linalg.norm(t - ts[f]) < 0.06
    o.close()


def test_ate_tool_alignment():
    # tools/evaluate_ate.py (TUM-style ATE for the trajectory the node writes, sdf_reconstruction.cpp:4-17)
    from tools import evaluate_ate as E
    rng = np.random.default_rng(0)
    gt = rng.normal(size=(50, 3))
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    est = (gt - 0.5) @ R.T
    assert E.ate_rmse(est, gt)[0] < 1e-12 and E.ate_rmse(est, gt, do_align=False)[0] > 0.5
    noisy = gt + 0.01
    assert E.ate_rmse(noisy, gt, do_align=False)[0] == pytest.approx(0.01 * np.sqrt(3), rel=1e-9)


def test_color_fusion_by_hand():
    # sdf.cpp:294-304: for a fronto-parallel plane n = (0,0,-1) the cosine is 1, so the colour weight is the
    # D/W weight and one observation stores the pixel's colour exactly; a second one is the running mean.
    o, depth = _single_voxel_case(0, 2.0)
    o.enable_color()
    CW, R, G, B = o.color()
    assert (CW == 0).all() and (R == np.float32(0.4)).all() and (G == np.float32(0.4)).all() and (B == np.float32(0.4)).all()   # sdf.cpp:30-34
    rgb = np.zeros((480, 640, 3), np.uint8); rgb[..., 0] = 200; rgb[..., 1] = 100; rgb[..., 2] = 50
    o.fuse_rgb(depth, rgb)
    CW, R, G, B = o.color()
    W = o.W
    upd = W > 0
    assert upd.any() and np.array_equal(CW[upd], W[upd])                  # cosine == 1
    assert (R[upd] == 200).all() and (G[upd] == 100).all() and (B[upd] == 50).all()
    assert (R[~upd] == np.float32(0.4)).all() and (CW[~upd] == 0).all()   # untouched voxels keep the initial grey
    rgb2 = np.zeros_like(rgb); rgb2[..., 0] = 100
    W1 = W.copy()
    o.fuse_rgb(depth, rgb2)
    CW, R, G, B = o.color()
    i = j = o.m // 2
    k = int(np.argmax(W1[i, j]))
    w1 = np.float32(W1[i, j, k]); w2 = np.float32(o.W[i, j, k] - w1)
    assert CW[i, j, k] == np.float32(w1 + w2)
    assert R[i, j, k] == np.float32(np.float32(np.float32(w1 * np.float32(200)) + np.float32(w2 * np.float32(100))) / np.float32(w1 + w2))
    assert G[i, j, k] == np.float32(np.float32(np.float32(w1 * np.float32(100)) + np.float32(w2 * np.float32(0))) / np.float32(w1 + w2))
    o.close()


def test_color_weight_is_cosine_of_the_normal(frames, K):
    # sdf.cpp:294,299: Color_W accumulates w * |n_z| / ||n||; with one observation Color_W / W = |n_z| of the pixel's
    # (unit) normal, so it lies in (0, 1] and is < 1 on slanted surfaces
    depth, Rs, ts = frames
    o = po.Oracle(m=64, use_coord_table=0); o.set_intrinsics(K); o.set_pose(Rs[0], ts[0])
    o.fuse_rgb(depth[0], np.full((480, 640, 3), 7, np.uint8))
    CW, R, G, B = o.color(); W = o.W
    upd = W > 0
    ratio = CW[upd] / W[upd]
    assert ratio.max() <= 1.0 + 1e-6 and ratio.min() >= 0.0 and (ratio < 0.9).any() and (ratio > 0.99).any()
    ok = CW > 0
    assert np.allclose(R[ok], 7, atol=1e-4) and np.allclose(B[ok], 7, atol=1e-4)    # a constant image stays constant
    o.close()


def test_interpolate_color_quirks():
    # sdf.cpp:164-217: world coordinates in; an exact voxel hit returns the raw 0..255 value (:193-198),
    # anything else is the inverse-L1 mean divided by 255 (:210-213); nothing in range -> 0/0 = NaN; alpha = 1
    o, depth = _single_voxel_case(0, 2.0)
    rgb = np.zeros((480, 640, 3), np.uint8); rgb[..., 0] = 200; rgb[..., 1] = 100; rgb[..., 2] = 50
    o.fuse_rgb(depth, rgb)
    CW = o.color()[0]
    i = j = o.m // 2
    k = int(np.argmax(CW[i, j]))
    assert CW[i, j, k] > 0 and CW[i, j, k + 1] > 0
    c = o.get_global_coordinates([i, j, k])
    out = o.interpolate_color([c])[0]
    assert tuple(out) == (200.0, 100.0, 50.0, 1.0)
    c2 = np.array(o.get_global_coordinates([i, j, k + 1]))
    mid = o.interpolate_color([(np.array(c) + c2) / 2 + [0.003, 0.002, 0.0]])[0]
    assert mid[0] == pytest.approx(200 / 255, rel=1e-5) and mid[1] == pytest.approx(100 / 255, rel=1e-5) and mid[3] == 1.0
    far = o.interpolate_color([[100.0, 100.0, 100.0]])[0]
    assert np.isnan(far[:3]).all() and far[3] == 1.0
    o.close()


def test_marching_cubes_sphere_is_closed_and_on_the_surface():
    # pcl::MarchingCubesSDF on an analytic sphere (create_circle fixture, sdf.cpp:62-93): vertices sit on the iso
    # surface (mesher frame: extent * index / m, i.e. world - origin - half a voxel, marching_cubes_sdf.cpp:123-125),
    # every edge is shared by exactly two triangles and V - E + F = 2
    m = 40
    o = po.Oracle(m=m, use_coord_table=0)
    o.create_circle(1.0, 0.0, 0.0, 1.25)
    (xyz,) = o.mesh(0.0)
    assert len(xyz) % 3 == 0 and len(xyz) > 900
    p = xyz.astype(np.float64) + [-3.0, -3.0, -0.5] + np.array([6.0, 6.0, 3.5]) / m / 2
    r = np.linalg.norm(p - [0.0, 0.0, 1.25], axis=1)
    assert abs(r.mean() - 1.0) < 0.02 and np.abs(r - 1.0).max() < 0.05
    # neighbouring cells compute a shared vertex from differently rounded corners (centre + step vs the next centre):
    # weld at 1e-4 m (voxels are 0.15 m here)
    uniq, inv = np.unique(np.rint(xyz.astype(np.float64) / 1e-4).astype(np.int64), axis=0, return_inverse=True)
    tri = inv.reshape(-1, 3)
    tri = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]        # drop degenerate slivers
    e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    ue, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    assert len(np.unique(tri)) - len(ue) + len(tri) == 2
    # iso level outside [0, 1): empty (marching_cubes_sdf.cpp:248-254); unseen voxels: no surface (:221-241)
    assert len(o.mesh(1.0)[0]) == 0 and len(o.mesh(-0.01)[0]) == 0
    o.W[:, :, : m // 2] = 0
    (half,) = o.mesh(0.0)
    assert 0 < len(half) < len(xyz)
    o.close()


def test_marching_cubes_single_cell_by_hand():
    # one corner below the iso level -> one triangle on the three edges that meet there, interpolated linearly
    m = 8
    o = po.Oracle(m=m, use_coord_table=0)
    o.W[...] = 1.0; o.D[...] = 1.0
    o.D[3, 3, 3] = -1.0                                       # corner 0 of cell (3,3,3), corner 1 of cell (2,3,3), ...
    (xyz,) = o.mesh(0.0)
    assert len(xyz) == 8 * 3                                  # the eight cells around the voxel, one triangle each
    vs = np.float32(6.0) / np.float32(m); vz = np.float32(3.5) / np.float32(m)
    c = np.array([np.float32(6.0) * np.float32(3) / np.float32(m)] * 2 + [np.float32(3.5) * np.float32(3) / np.float32(m)], np.float32)
    # cell (3,3,3): configuration 1 -> edges 0 (x), 8 (y), 3 (z, from corner 3 to corner 0): mu = 0.5 on each
    cell = xyz[-3:]
    expect = {(float(c[0] + np.float32(0.5) * vs), float(c[1]), float(c[2])),
              (float(c[0]), float(c[1] + np.float32(0.5) * vs), float(c[2])),
              (float(c[0]), float(c[1]), float(np.float32(c[2] + vz) + np.float32(0.5) * np.float32(c[2] - np.float32(c[2] + vz))))}
    assert {tuple(map(float, v)) for v in cell} == expect
    o.close()
