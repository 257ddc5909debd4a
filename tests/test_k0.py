"""K0 — the pre-processing the reference's node applies before the boundary (sdf_reconstruction.cpp:37-49:
pcl::FastBilateralFilter with PCL's defaults, pcl::IntegralImageNormalEstimation AVERAGE_3D_GRADIENT / 0.02 / 10).
PCL is un-vendored, so parity is UNPINNED at that boundary; these are known-answer tests of the oracle's definition
(oracle.cpp: k0_bilateral, k0_normals) from first principles.  tests/test_gpu_k0.py holds the device to it bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tools import synth

H, W = 480, 640
K = synth.K_DEFAULT


@pytest.fixture(scope="module")
def orc():
    o = po.Oracle(m=16, use_coord_table=0); o.set_intrinsics(K)
    yield o
    o.close()


def plane_depth(n, d):
    """depth image of the plane n . p = d seen from the origin"""
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    ray = np.stack([(u - K[2]) / K[0], (v - K[5]) / K[4], np.ones_like(u)], -1)
    return (d / (ray @ np.asarray(n, float))).astype(np.float32)


def test_constant_depth_is_a_fixed_point(orc):
    d = np.full((H, W), 2.0, np.float32)
    zf, n = orc.preprocess(d)
    assert np.abs(zf - 2.0).max() < 1e-6                      # one occupied depth bin: sum / count = z
    inner = n[12:-12, 12:-12]
    assert np.isfinite(inner).all()
    assert np.abs(inner - np.array([0, 0, -1.0])).max() < 1e-6   # fronto-parallel plane, normal faces the camera


def test_tilted_plane_normal_and_smoothness(orc):
    nrm = np.array([0.3, -0.2, -1.0]); nrm /= np.linalg.norm(nrm)
    d = plane_depth(nrm, -2.0)                                # n . p = -2 with n_z < 0: in front of the camera
    assert d.min() > 1.0 and d.max() < 4.0
    zf, n = orc.preprocess(d)
    assert np.abs(zf - d)[20:-20, 20:-20].max() < 1.5e-2      # bilateral-grid quantisation on a slope: < 1.5 cm (sigma_r = 5 cm)
    inner = n[20:-20, 20:-20]
    assert np.isfinite(inner).all()
    assert (inner @ nrm).min() > 0.99                         # ripples of the grid quantisation: a few degrees at most
    exact = orc.k0_normals(d)[20:-20, 20:-20]                 # the normal stage alone on the exact plane
    assert np.isfinite(exact).all() and (exact @ nrm).min() > 0.999999     # averaged 3-D gradients of a plane: its normal
    assert (np.einsum("ijk,ijk->ij", inner, np.dstack([np.zeros_like(zf), np.zeros_like(zf), zf])[20:-20, 20:-20]) < 0).all()


def test_step_edge_is_preserved_and_has_no_normals(orc):
    d = np.full((H, W), 1.5, np.float32)
    d[:, 320:] = 2.5                                          # a 1 m step = 20 sigma_r: no mixing across it
    zf, n = orc.preprocess(d)
    assert np.abs(zf[:, :300] - 1.5).max() < 1e-4 and np.abs(zf[:, 340:] - 2.5).max() < 1e-4
    assert np.abs(zf[:, 318:322] - d[:, 318:322]).max() < 1e-3
    assert np.isnan(n[100, 318:322]).all()                    # window size = distance to the discontinuity <= 2: no normal
    assert np.isfinite(n[100, 300]).all() and np.isfinite(n[100, 340]).all()
    # window grows with the distance from the edge (capped at 10): still a valid, correct normal
    assert np.abs(n[100, 312] - np.array([0, 0, -1.0])).max() < 1e-6


def test_invalid_pixels_stay_invalid_and_are_discontinuities(orc):
    d = np.full((H, W), 2.0, np.float32)
    d[200:210, 300:330] = np.nan; d[50, 60] = 0.0; d[51, 61] = np.inf
    zf, n = orc.preprocess(d)
    assert np.isnan(zf[200:210, 300:330]).all() and np.isnan(zf[50, 60]) and np.isnan(zf[51, 61])
    assert np.isfinite(zf[np.isfinite(d) & (d > 0)]).all()
    assert np.isnan(n[205, 298]).all() and np.isfinite(n[205, 280]).all()


def test_filter_denoises_and_normals_beat_k1_on_noisy_depth(orc, frames):
    depth, Rs, ts = frames
    clean = depth[0]
    noisy = synth.add_sensor_noise(clean, seed=1234, dropout=0.0)
    zf, n0 = orc.preprocess(noisy)
    _, n_clean = orc.backproject(clean)                       # K1's normals on the noise-free frame = ground truth
    _, n1 = orc.backproject(noisy)                            # K1's 4-neighbour normals on the noisy frame
    flat = (np.abs(np.gradient(clean)[0]) + np.abs(np.gradient(clean)[1]) < 0.01)[1:-1, 1:-1]

    def roughness(x):                                         # pixel-to-pixel noise: |x - mean of the 4 neighbours|
        e = x - clean
        return np.abs(e[1:-1, 1:-1] - 0.25 * (e[:-2, 1:-1] + e[2:, 1:-1] + e[1:-1, :-2] + e[1:-1, 2:]))[flat].mean()
    # the grid's quantisation adds a smooth bias on slopes (< 1.5 cm, see the tilted-plane test), so the filter is
    # judged on what it is for: the pixel-to-pixel noise that wrecks finite-difference normals
    assert roughness(zf) < 0.7 * roughness(noisy), (roughness(zf), roughness(noisy))      # measured 0.55
    ok = np.isfinite(n0[..., 0]) & np.isfinite(n1[..., 0]) & np.isfinite(n_clean[..., 0])
    c0 = (n0[ok] * n_clean[ok]).sum(-1).mean(); c1 = (n1[ok] * n_clean[ok]).sum(-1).mean()
    assert c0 > 0.97 and c1 < 0.8, (c0, c1)                   # measured 0.984 vs 0.62


def test_oracle_pipeline_uses_k0_when_asked(frames):
    depth, Rs, ts = frames
    noisy = synth.add_sensor_noise(depth[:3], seed=7)
    a = po.Oracle(m=32, use_coord_table=0, preprocess=1); b = po.Oracle(m=32, use_coord_table=0)
    for o in (a, b):
        o.set_intrinsics(K); o.set_pose(Rs[0], ts[0])
    na = a.fuse(noisy[0]); nb = b.fuse(noisy[0])
    assert na > 1000 and nb > 1000 and not np.array_equal(a.D, b.D)
    a.close(); b.close()
