"""The oracle against the committed golden fixtures (tests/golden, made by make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tools import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "track_fuse_m32.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_synthetic_frames_are_reproducible(gold, frames):
    depth, Rs, ts = frames
    assert np.array_equal(depth[:6, ::60, ::80], gold["depth_probe"])
    sha = np.frombuffer(hashlib.sha256(depth[:6].tobytes()).digest(), np.uint8)
    assert np.array_equal(sha, gold["depth_sha256"])
    assert np.array_equal(Rs[:6], gold["R_gt"]) and np.array_equal(ts[:6], gold["t_gt"])
    assert np.isfinite(depth).all() and depth.min() > 0.3 and depth.max() < 7.0


@pytest.mark.parametrize("metric", [0, 1])
def test_oracle_reproduces_golden(metric, gold, frames):
    depth, Rs, ts = frames
    g = lambda k: gold["m%d_%s" % (metric, k)]
    o = po.Oracle(m=32, use_coord_table=0, metric=metric, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    o.set_intrinsics(gold["K"])
    for f in range(3):
        o.set_pose(Rs[f], ts[f])
        assert o.fuse(depth[f]) == g("n_updated")[f]
    assert np.array_equal(o.D, g("D")) and np.array_equal(o.W, g("W"))
    o.set_pose(Rs[3], ts[3])
    A, b, st = o.linearize(depth[3])
    assert np.array_equal(A, g("A")) and np.array_equal(b, g("b"))
    J, psi, flag = o.linearize_pixels(depth[3])
    assert np.array_equal(flag, g("flag")) and np.array_equal(J[::16], g("J")) and np.array_equal(psi[::16], g("psi"))
    assert np.array_equal(np.frombuffer(hashlib.sha256(J.tobytes() + psi.tobytes()).digest(), np.uint8), g("J_sha256"))
    o.set_pose(Rs[2], ts[2])
    o.track(depth[3])
    R, t = o.get_pose()
    assert np.abs(R - g("R_tracked")).max() < 1e-13 and np.abs(t - g("t_tracked")).max() < 1e-13
    v, ok = o.interpolate_distance(g("sample_pts"))
    assert np.array_equal(ok, g("sample_ok")) and np.array_equal(v, g("sample_val"), equal_nan=True)
    o.close()


def test_exp_map_golden(gold):
    for tw, R, t in zip(gold["exp_twist"], gold["exp_R"], gold["exp_t"]):
        Ro, to = po.exp_map(tw)
        assert np.abs(Ro - R).max() < 1e-15 and np.abs(to - t).max() < 1e-15


GOLD_CM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "color_mesh_m32.npz")


def test_oracle_reproduces_color_and_mesh_golden(frames):
    """Colour fusion, colour sampling and the mesher against tests/golden/color_mesh_m32.npz."""
    gold = np.load(GOLD_CM)
    depth, Rs, ts = frames
    o = po.Oracle(m=32, use_coord_table=0, metric=0); o.set_intrinsics(synth.K_DEFAULT)
    sha = hashlib.sha256()
    for f in range(3):
        rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
        sha.update(rgb.tobytes())
        o.set_pose(Rs[f], ts[f])
        assert o.fuse_rgb(depth[f], rgb) == gold["n_updated"][f]
    assert np.array_equal(np.frombuffer(sha.digest(), np.uint8), gold["rgb_sha256"])     # the synthetic colour images are reproducible
    for a, k in zip(o.color(), ("Color_W", "R", "G", "B")):
        assert np.array_equal(a, gold[k], equal_nan=True), k
    assert np.array_equal(o.interpolate_color(gold["sample_pts"]), gold["sample_rgba"], equal_nan=True)
    for iso, tag in ((0.0, "iso0"), (0.1, "iso01")):
        xyz, world, rgba = o.mesh(iso, world=True, colors=True)
        assert np.array_equal(xyz, gold["mesh_%s_xyz" % tag]) and np.array_equal(world, gold["mesh_%s_world" % tag])
        assert np.array_equal(rgba, gold["mesh_%s_rgba" % tag], equal_nan=True)
    o.close()
