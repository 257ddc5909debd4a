"""GPU parity of the mesher (pcl::MarchingCubesSDF::performReconstruction, marching_cubes_sdf.cpp:243-287, and the
marker post-processing of SDF::visualize, sdf.cpp:352-385) against the CPU oracle, through the C ABI.
fp32 arithmetic restated operation for operation: vertices, their order and their colours are bit-exact."""
import numpy as np
import pytest

import tracking_sdf_b200 as T
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.gpu


def fused_pair(m, K, frames, n, color=True):
    depth, Rs, ts = frames
    o = po.Oracle(m=m, use_coord_table=0, metric=0); o.set_intrinsics(K)
    g = T.Tsdf(T.default_config(m=m, metric=0)); g.set_intrinsics(K)
    for f in range(n):
        o.set_pose(Rs[f], ts[f])
        if color:
            rgb = synth.synth_rgb(depth[f], Rs[f], ts[f])
            o.fuse_rgb(depth[f], rgb); g.fuse_rgb(depth[f], rgb, Rs[f], ts[f])
        else:
            o.fuse(depth[f]); g.fuse(depth[f], Rs[f], ts[f])
    return o, g


@pytest.mark.parametrize("m", [64, 96, 100])
def test_mesh_bit_exact(gpu_lib, frames, K, m):
    o, g = fused_pair(m, K, frames, 3)
    xo, wo, co = o.mesh(0.0, world=True, colors=True)
    xg, wg, cg = g.mesh(0.0, world=True, colors=True)
    assert len(xo) > 3000 and len(xo) % 3 == 0
    assert xo.shape == xg.shape and np.array_equal(xo, xg)              # same triangles, same order, same bits
    assert np.array_equal(wo, wg) and np.array_equal(co, cg, equal_nan=True)
    # other iso levels inside [0, 1); outside -> empty (marching_cubes_sdf.cpp:248-254)
    for iso in (0.05, 0.2):
        assert np.array_equal(o.mesh(iso)[0], g.mesh(iso)[0])
    assert len(g.mesh(-0.1)[0]) == 0 and len(g.mesh(1.0)[0]) == 0 and len(o.mesh(1.0)[0]) == 0
    g.close(); o.close()


def test_mesh_list_overflow_falls_back_to_two_sweeps(gpu_lib, frames, K, monkeypatch):
    """The one-sweep mesher appends surface cells to a compact list; when the list is too small the extraction falls
    back to the second sweep over the store and grows the list for the next call.  Both paths give the oracle's mesh."""
    monkeypatch.setenv("TSDF_B200_MC_CELLS", "64")
    o, g = fused_pair(64, K, frames, 3, color=False)
    xo = o.mesh(0.0)[0]
    x1 = g.mesh(0.0)[0]                      # list of 64 cells overflows: two-sweep fallback
    x2 = g.mesh(0.0)[0]                      # list grown: one sweep + list emit
    assert len(xo) > 3000 and np.array_equal(xo, x1) and np.array_equal(xo, x2)
    g.close(); o.close()


def test_mesh_of_empty_and_partially_seen_volume(gpu_lib, frames, K):
    depth, Rs, ts = frames
    g = T.Tsdf(T.default_config(m=32)); g.set_intrinsics(K)
    assert len(g.mesh()[0]) == 0                                        # nothing observed: W = 0 everywhere
    o = po.Oracle(m=32, use_coord_table=0); o.set_intrinsics(K)
    d = depth[0].copy(); d[:, 320:] = np.nan                            # half the image: cells straddling seen/unseen are skipped (:221)
    o.set_pose(Rs[0], ts[0]); o.fuse(d); g.fuse(d, Rs[0], ts[0])
    xo = o.mesh()[0]; xg = g.mesh()[0]
    assert len(xo) > 0 and np.array_equal(xo, xg)
    with pytest.raises(T.TsdfError):
        g.mesh(colors=True)                                             # no colour store
    g.close(); o.close()


def test_mesh_after_upload_of_an_analytic_sphere(gpu_lib, K):
    # create_circle-style fixture (sdf.cpp:62-93 idea): D = distance to a sphere, W = 1 -> closed surface
    m = 48
    o = po.Oracle(m=m, use_coord_table=0)
    o.create_circle(1.0, 0.0, 0.0, 1.25)
    g = T.Tsdf(T.default_config(m=m)); g.upload(o.D, o.W)
    xo = o.mesh()[0]; xg = g.mesh()[0]
    assert len(xo) > 1000 and np.array_equal(xo, xg)
    # vertices lie on the sphere (mesher frame = world - origin, minus the half-voxel shift of :123-125)
    p = xg.astype(np.float64) + [-3.0, -3.0, -0.5] + np.array([6.0, 6.0, 3.5]) / m / 2
    r = np.linalg.norm(p - [0.0, 0.0, 1.25], axis=1)
    assert abs(r.mean() - 1.0) < 0.02 and r.std() < 0.02
    g.close(); o.close()


def test_color_and_mesh_against_committed_golden(gpu_lib, frames, K):
    """The CUDA path against tests/golden/color_mesh_m32.npz (made by the oracle, tests/golden/make_golden.py)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "color_mesh_m32.npz"))
    depth, Rs, ts = frames
    g = T.Tsdf(T.default_config(m=32, metric=0)); g.set_intrinsics(K)
    for f in range(3):
        assert g.fuse_rgb(depth[f], synth.synth_rgb(depth[f], Rs[f], ts[f]), Rs[f], ts[f]) == gold["n_updated"][f]
    for a, k in zip(g.download_color(), ("Color_W", "R", "G", "B")):
        assert np.array_equal(a, gold[k], equal_nan=True), k
    assert np.array_equal(g.interpolate_color(gold["sample_pts"]), gold["sample_rgba"], equal_nan=True)
    for iso, tag in ((0.0, "iso0"), (0.1, "iso01")):
        xyz, world, rgba = g.mesh(iso, world=True, colors=True)
        assert np.array_equal(xyz, gold["mesh_%s_xyz" % tag]) and np.array_equal(world, gold["mesh_%s_world" % tag])
        assert np.array_equal(rgba, gold["mesh_%s_rgba" % tag], equal_nan=True)
    g.close()
