"""THE PIN: oracle/oracle.cpp (the restatement every GPU parity test checks against) versus THE REFERENCE ITSELF —
the four hot-path translation units of /root/reference compiled unmodified against oracle/shim/
(oracle/Makefile `ref` -> oracle/_ref/libtsdf_ref.so, bound by oracle/pyref.py).

Bit-equality everywhere the reference's own arithmetic decides the result.  The two places that depend on Eigen
internals the shim can only state, not verify (the 6x6 `A.inverse()*b`, camera_tracking.cpp:191, and the SVD
inside `Affine3d::rotation()`, :237-238) are compared at a few ulp and say so.

Needs either /root/reference (builds _ref) or a prebuilt oracle/_ref/libtsdf_ref.so; skipped otherwise (the GPU
box has neither the sources nor a reason to re-run this: the committed goldens carry the reference's results).
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref as pr
from tools import synth

pytestmark = pytest.mark.skipif(not pr.available(), reason="no /root/reference and no prebuilt oracle/_ref")

M = 48
KW = dict(m=M, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))


def pair(**kw):
    a = dict(KW); a.update(kw)
    o = po.Oracle(use_coord_table=0, **a); r = pr.Reference(**a)
    o.set_intrinsics(synth.K_DEFAULT); r.set_intrinsics(synth.K_DEFAULT)
    return o, r


def ragged(depth, seed):
    """Invalid pixels as a sensor produces them: NaN speckle, a NaN block, zeros, an infinity."""
    d = depth.copy()
    rng = np.random.default_rng(seed)
    d[rng.random(d.shape) < 0.05] = np.nan
    d[100:160, 200:330] = np.nan
    d[300:310, 10:50] = 0.0
    d[7, 9] = np.inf
    return d


def far_block(depth):
    """A block of pixels whose depth (9 m) puts the back-projected points outside the 6 x 6 x 3.5 m volume."""
    d = depth.copy()
    d[200:260, 300:420] = 9.0
    return d


@pytest.fixture(scope="module")
def fused(frames):
    depth, Rs, ts = frames
    o, r = pair()
    for f in range(3):
        o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
        o.fuse(depth[f]); r.fuse(depth[f])
    yield o, r
    o.close(); r.close()


def test_constants_and_maps():
    # sdf.cpp:19-21, camera_tracking.cpp:13-17, sdf.h:113-157
    for m in (32, 100, 256):
        o = po.Oracle(m=m, use_coord_table=0); r = pr.Reference(m=m)
        assert np.array_equal(o.constants(), r.constants())
        rng = np.random.default_rng(m)
        for _ in range(100):
            ijk = rng.integers(-2, m + 2, 3)
            assert o.get_array_index(*map(int, ijk)) == r.get_array_index(*map(int, ijk))
            ijk = np.clip(ijk, 0, m - 1)
            g = o.get_global_coordinates(ijk)
            assert np.array_equal(g, r.get_global_coordinates(ijk))
            w = g + rng.normal(0, 0.3, 3)
            assert np.array_equal(o.get_voxel_coordinates(w), r.get_voxel_coordinates(w))
            idx = int(rng.integers(0, m ** 3))
            assert np.array_equal(o.get_voxel_coordinates_idx(idx), r.get_voxel_coordinates_idx(idx))
        assert (r.D == np.float32(15.5)).all() and (r.W == 0).all()           # sdf.cpp:29-31
        assert np.array_equal(o.get_pose()[0], r.get_pose()[0]) and np.array_equal(o.get_pose()[1], r.get_pose()[1])   # camera_tracking.cpp:5-8
        o.close(); r.close()


def test_set_camera_transformation(frames):
    # camera_tracking.cpp:59-65: Matrix3d::inverse() and -1 * (rot_inv * trans), incl. non-orthonormal input
    _, Rs, ts = frames
    o, r = pair(m=8)
    rng = np.random.default_rng(3)
    for f in range(12):
        R = Rs[f] if f < 8 else Rs[f] + rng.normal(0, 0.05, (3, 3))
        o.set_pose(R, ts[f]); r.set_pose(R, ts[f])
        for a, b in zip(o.get_pose_inv(), r.get_pose_inv()):
            assert np.array_equal(a, b)
    o.close(); r.close()


def test_update_three_frames(fused):
    # SDF::update sdf.cpp:224-292 — D and W bit-equal after three fused frames
    o, r = fused
    assert np.array_equal(o.D, r.D) and np.array_equal(o.W, r.W)
    assert (o.W > 0).sum() > 1000


def test_update_ragged_and_counts(frames):
    depth, Rs, ts = frames
    o, r = pair()
    for f in range(3):
        d = ragged(depth[f], f)
        o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
        assert o.fuse(d) == r.fuse(d)
    assert np.array_equal(o.D, r.D) and np.array_equal(o.W, r.W)
    o.close(); r.close()


def test_update_without_intrinsics_is_refused():
    # sdf.cpp:227-229: the reference exit(0)s; bridge and oracle both report -1 instead
    o = po.Oracle(m=8, use_coord_table=0); r = pr.Reference(m=8)
    d = np.ones((480, 640), np.float32)
    assert o.fuse(d) == -1 and r.fuse(d) == -1
    o.close(); r.close()


def test_interpolate_distance(fused):
    # sdf.cpp:127-163 incl. the early return, (int) truncation, out-of-range neighbours, NaN / huge inputs
    o, r = fused
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1.5, M + 0.5, (40000, 3))
    ii = np.argwhere(o.W > 0)
    centres = ii[rng.integers(0, len(ii), 2000)].astype(np.float64)
    near = centres + rng.uniform(-1, 1, centres.shape)
    special = np.array([[np.nan, 1, 1], [1e12, 1, 1], [-1e12, 2, 2], [3e9, 3e9, 3e9], [-0.5, -0.5, -0.5], [-0.999, 5, 5],
                        [M - 1, M - 1, M - 1], [M - 0.5, 3, 3], [np.inf, 0, 0], [0, 0, 0], [M - 1.000001, 1, 1]])
    allp = np.concatenate([pts, centres, near, special])
    vo, oko = o.interpolate_distance(allp); vr, okr = r.interpolate_distance(allp)
    assert np.array_equal(oko, okr)
    assert np.array_equal(vo, vr, equal_nan=True)
    assert oko.sum() > 3000 and np.isnan(vo).sum() > 100


def test_normal_equations_single_thread(fused, frames):
    # camera_tracking.cpp:79-189 with one OpenMP thread = the oracle's canonical order: A and b bit-equal
    depth, Rs, ts = frames
    o, r = fused
    pr.set_num_threads(1)
    try:
        for f in (3, 4):
            for d in (depth[f], ragged(depth[f], 10 + f)):
                o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
                A_o, b_o, st = o.linearize(d); A_r, b_r = r.linearize(d)
                assert st["n_oob"] == 0 and st["n_valid"] > 10000
                assert np.array_equal(A_o, A_r) and np.array_equal(b_o, b_r)
    finally:
        pr.set_num_threads(po.num_threads())


def test_normal_equations_threaded(fused, frames):
    # the reference's own result depends on its thread count (per-thread partial sums, :146-189): order only
    depth, Rs, ts = frames
    o, r = fused
    pr.set_num_threads(4)
    try:
        o.set_pose(Rs[3], ts[3]); r.set_pose(Rs[3], ts[3])
        A_o, b_o, _ = o.linearize(depth[3]); A_r, b_r = r.linearize(depth[3])
        assert np.abs(A_o - A_r).max() <= 1e-13 * np.abs(A_o).max() and np.abs(b_o - b_r).max() <= 1e-13 * np.abs(b_o).max()
    finally:
        pr.set_num_threads(po.num_threads())


def test_partial_derivative_per_pixel(fused, frames):
    # CameraTracking::get_partial_derivative camera_tracking.cpp:246-363: J (6), psi and the outcome per pixel
    depth, Rs, ts = frames
    o, r = fused
    for f, d in ((3, depth[3]), (5, ragged(depth[5], 2))):
        o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
        Jo, po_, fo = o.linearize_pixels(d); Jr, pr_, fr = r.linearize_pixels(d)
        assert np.array_equal(fo, fr)
        assert np.array_equal(Jo, Jr) and np.array_equal(po_, pr_)
        assert (fo == 1).sum() > 10000 and (fo == 3).sum() > 100


def test_out_of_volume_pixels_flagged_alike(fused, frames):
    # :261-268 — a pose that pushes part of the cloud out of the volume: same pixels are out for both
    depth, Rs, ts = frames
    o, r = fused
    d = far_block(depth[3])
    o.set_pose(Rs[3], ts[3]); r.set_pose(Rs[3], ts[3])
    Jo, po_, fo = o.linearize_pixels(d); Jr, pr_, fr = r.linearize_pixels(d)
    assert (fo == 2).sum() > 100 and (fo == 1).sum() > 10000 and np.array_equal(fo, fr) and np.array_equal(Jo, Jr)


def test_trap5_stale_state_is_the_references_not_the_oracles(fused, frames):
    """SURVEY TRAP 5, now demonstrated on the reference's own code: an out-of-volume centre returns early
    (camera_tracking.cpp:261-268) leaving is_interpolated / SDF_derivative / int_dist at the PREVIOUS pixel's
    values, and the caller (:178-182) adds that pixel again.  With one thread the order is fixed, so the
    reference's A equals the oracle's per-pixel records summed with exactly that rule; the oracle proper treats
    such pixels as invalid — on benchmark inputs n_oob = 0 and the two coincide."""
    depth, Rs, ts = frames
    o, r = fused
    d = far_block(depth[3])
    o.set_pose(Rs[3], ts[3]); r.set_pose(Rs[3], ts[3])
    pr.set_num_threads(1)
    try:
        A_r, b_r = r.linearize(d)
    finally:
        pr.set_num_threads(po.num_threads())
    A_o, b_o, st = o.linearize(d)
    J, psi, flag = o.linearize_pixels(d)
    assert st["n_oob"] > 100 and not np.array_equal(A_o, A_r)
    A = np.zeros((6, 6)); b = np.zeros(6)
    live = False; Jp = np.zeros(6); pp = 0.0
    for p in range(len(flag)):                       # reference loop order: i outer, j inner = record order
        if flag[p] == 0:                             # NaN point: `continue` before the call, state untouched
            continue
        if flag[p] == 1:
            live = True; Jp = J[p].astype(np.float64); pp = float(psi[p])
        elif flag[p] == 3:                           # a sample failed: is_interpolated = false
            live = False
        if live:                                     # flag 2 (out of volume) re-adds the previous pixel
            A += np.outer(Jp, Jp); b += pp * Jp
    assert np.array_equal(A, A_r) and np.array_equal(b, b_r)


def test_solve_and_pose_update(fused, frames):
    # camera_tracking.cpp:191-192, 237-239.  A.inverse()*b goes through Eigen's 6x6 PartialPivLU inverse, whose
    # blocked triangular solves the shim can only state (eigen_shim.h): compared at a few ulp, not bitwise.
    depth, Rs, ts = frames
    o, r = fused
    o.set_pose(Rs[3], ts[3]); r.set_pose(Rs[3], ts[3])
    A, b, _ = o.linearize(depth[3])
    tw_o, sing = o.apply_update(A, b); tw_r = r.apply_update(A, b)
    assert sing == 0
    assert np.abs(tw_o - tw_r).max() <= 1e-14 * np.abs(tw_o).max()
    (Ro, to), (Rr, tr) = o.get_pose(), r.get_pose()
    assert np.abs(Ro - Rr).max() <= 1e-15 and np.abs(to - tr).max() <= 1e-15
    # same twist in -> same pose out, bitwise: the exp map and the update expressions themselves are pinned
    o.set_pose(Rs[3], ts[3]); r.set_pose(Rs[3], ts[3])
    Ainv = np.linalg.inv(A)
    tw_o, _ = o.apply_update(np.eye(6), Ainv @ b); tw_r = r.apply_update(np.eye(6), Ainv @ b)
    assert np.array_equal(tw_o, tw_r)
    (Ro, to), (Rr, tr) = o.get_pose(), r.get_pose()
    assert np.array_equal(Ro, Rr) and np.array_equal(to, tr)
    for a, b2 in zip(o.get_pose_inv(), r.get_pose_inv()):
        assert np.array_equal(a, b2)


def test_estimate_new_position(fused, frames):
    # camera_tracking.cpp:66-245, 10 fixed iterations from the previous frame's pose, then the reference default
    # (20 iterations, signed stop test :216-224)
    depth, Rs, ts = frames
    o, r = fused
    pr.set_num_threads(1)
    try:
        o.set_pose(Rs[2], ts[2]); r.set_pose(Rs[2], ts[2])
        so = o.track(depth[3]); sr = r.track(depth[3])
        assert so["iterations"] == sr["iterations"] == 10
        (Ro, to), (Rr, tr) = o.get_pose(), r.get_pose()
        assert np.abs(Ro - Rr).max() <= 1e-13 and np.abs(to - tr).max() <= 1e-13
        assert np.linalg.norm(to - ts[3]) < 0.08           # m = 48: 12.5 cm voxels
    finally:
        pr.set_num_threads(po.num_threads())
    o2, r2 = pair(gauss_newton_max_iteration=20, maximum_twist_diff=0.001)
    for f in range(2):
        o2.set_pose(Rs[f], ts[f]); r2.set_pose(Rs[f], ts[f]); o2.fuse(depth[f]); r2.fuse(depth[f])
    so = o2.track(depth[2]); sr = r2.track(depth[2])
    assert so["iterations"] == sr["iterations"] and so["stopped"] == sr["stopped"]
    (Ro, to), (Rr, tr) = o2.get_pose(), r2.get_pose()
    assert np.abs(Ro - Rr).max() <= 1e-12 and np.abs(to - tr).max() <= 1e-12
    o2.close(); r2.close()


def test_trap12_singular_system():
    # camera_tracking.cpp:191 is unguarded: an empty volume gives A = 0 and the reference's pose becomes NaN;
    # the oracle keeps the pose and reports `singular`
    o, r = pair(m=16)
    d = np.full((480, 640), 2.0, np.float32)
    R0, t0 = o.get_pose()
    so = o.track(d); sr = r.track(d)
    assert so["singular"] == 1 and sr["singular"] == 1
    assert np.array_equal(o.get_pose()[1], t0) and not np.isfinite(r.get_pose()[1]).all()
    o.close(); r.close()


def test_exp_map():
    # eigen_utils.cpp:40-128 incl. both small-angle guards
    rng = np.random.default_rng(11)
    tw = rng.uniform(-0.5, 0.5, (64, 6)); tw[0] = 0; tw[1, 3:] = 1e-5; tw[2, 3:] = 1e-9; tw[3, 3:] = [2.4e-4, 0, 0]; tw[4, 3:] = [0, 2.6e-4, 0]
    tw[5] = [1, 2, 3, np.pi / 2, 0, 0]
    for x in tw:
        (Ro, to), (Rr, tr) = po.exp_map(x), pr.exp_map(x)
        assert np.array_equal(Ro, Rr) and np.array_equal(to, tr)


def test_create_circle_and_gradient():
    # sdf.cpp:99-126 (the reference's own "testing issues" helper) + the Jacobian on it
    o, r = pair(m=40)
    o.create_circle(1.0, 0.2, -0.3, 1.2); r.create_circle(1.0, 0.2, -0.3, 1.2)
    assert np.array_equal(o.D, r.D) and np.array_equal(o.W, r.W)
    d = synth.render_sequence(1)[0][0]
    Jo, po_, fo = o.linearize_pixels(d); Jr, pr_, fr = r.linearize_pixels(d)
    assert np.array_equal(fo, fr) and np.array_equal(Jo, Jr) and np.array_equal(po_, pr_)
    o.close(); r.close()


def test_color_fusion_and_sampling(frames):
    # sdf.cpp:294-304 (cosine-weighted colour mean) and SDF::interpolate_color sdf.cpp:164-217
    depth, Rs, ts = frames
    o, r = pair()
    for f in range(3):
        d = depth[f] if f < 2 else ragged(depth[f], 1)
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
        assert o.fuse_rgb(d, rgb) == r.fuse_rgb(d, rgb)
    assert np.array_equal(o.D, r.D) and np.array_equal(o.W, r.W)
    for a, b in zip(o.color(), r.color()):
        assert np.array_equal(a, b)
    rng = np.random.default_rng(17)
    pts = rng.uniform([-3.1, -3.1, -0.6], [3.1, 3.1, 3.1], (20000, 3))
    ii = np.argwhere(o.color()[0] > 0)
    centres = np.stack([o.get_global_coordinates(q) for q in ii[rng.integers(0, len(ii), 300)]])
    allp = np.concatenate([pts, centres, centres + rng.normal(0, 0.05, centres.shape)])
    co, cr = o.interpolate_color(allp), r.interpolate_color(allp)
    assert np.array_equal(co, cr, equal_nan=True)
    assert np.isfinite(co[:, 0]).sum() > 500
    o.close(); r.close()


@pytest.mark.parametrize("iso", [0.0, 0.1])
def test_marching_cubes(frames, iso):
    # pcl::MarchingCubesSDF::performReconstruction marching_cubes_sdf.cpp:243-287 (+ createSurface, getNeighborList1D)
    depth, Rs, ts = frames
    o, r = pair()
    for f in range(3):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        o.set_pose(Rs[f], ts[f]); r.set_pose(Rs[f], ts[f])
        o.fuse_rgb(depth[f], rgb); r.fuse_rgb(depth[f], rgb)
    xyz_o, world_o, rgba_o = o.mesh(iso, world=True, colors=True)
    xyz_r = r.mesh(iso)
    assert len(xyz_o) > 300 and xyz_o.shape == xyz_r.shape and np.array_equal(xyz_o, xyz_r)
    if iso == 0.0:
        # one pass of SDF::visualize's loop (sdf.cpp:324-389): marker points and per-vertex colours
        world_r, rgba_r = r.visualize()
        assert np.array_equal(world_o, world_r) and np.array_equal(rgba_o, rgba_r, equal_nan=True)
    assert len(r.mesh(1.5)) == 0 and len(o.mesh(1.5)[0]) == 0     # :248-254 iso level outside [0,1)
    o.close(); r.close()


def test_closed_loop_20_frames():
    """20 frames of the node's loop (sdf_reconstruction.cpp:69-74: track, then update) run by the reference and
    by the oracle, free-running, one thread so that the summation order is the canonical one."""
    depth, Rs, ts = synth.render_sequence(20)
    o, r = pair(m=64)
    pr.set_num_threads(1)
    try:
        o.set_pose(Rs[0], ts[0]); r.set_pose(Rs[0], ts[0])
        o.fuse(depth[0]); r.fuse(depth[0])
        worst = 0.0
        for f in range(1, 20):
            o.track(depth[f]); r.track(depth[f])
            (Ro, to), (Rr, tr) = o.get_pose(), r.get_pose()
            worst = max(worst, np.abs(Ro - Rr).max(), np.abs(to - tr).max())
            o.fuse(depth[f]); r.fuse(depth[f])
        assert worst <= 1e-11, worst
        bad = (o.D != r.D) | (o.W != r.W)
        assert bad.mean() <= 1e-5, bad.sum()       # ulp-level pose differences may flip single (int) truncations
        assert np.linalg.norm(o.get_pose()[1] - ts[19]) < 0.25   # sanity only: 9 cm voxels at m = 64
    finally:
        pr.set_num_threads(po.num_threads())
    o.close(); r.close()
