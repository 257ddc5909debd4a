"""Synthetic inputs: fr1/plant camera path + analytic-scene depth frames (host tool)."""
import ctypes
import os

import numpy as np

from . import build as _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAJ = os.path.join(ROOT, "data", "fr1_plant_gt_every4.txt")

# ROS/TUM default intrinsics: fx = fy = 525, cx = 319.5, cy = 239.5
K_DEFAULT = np.array([525.0, 0.0, 319.5, 0.0, 525.0, 239.5, 0.0, 0.0, 1.0])
WIDTH, HEIGHT = 640, 480

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build_synth())
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.synth_render_depth.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        _lib.synth_render_depth.restype = None
        _lib.synth_clearance.argtypes = [dp]
        _lib.synth_clearance.restype = ctypes.c_double
    return _lib


def quat_to_rot(q):
    """(qx,qy,qz,qw) -> 3x3 rotation, camera optical frame -> world."""
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def load_trajectory(path=TRAJ):
    """-> (stamps [n], R [n,3,3], t [n,3]); pose = camera->world (x_w = R p + t)."""
    a = np.loadtxt(path, comments="#")
    stamps = a[:, 0]
    t = a[:, 1:4].copy()
    R = np.stack([quat_to_rot(q / np.linalg.norm(q)) for q in a[:, 4:8]])
    return stamps, R, t


def frame_index(i, n):
    """Ping-pong index so that arbitrarily long runs stay on the (continuous) path."""
    period = 2 * (n - 1)
    j = i % period
    return j if j < n else period - j


def render_depth(R, t, K=K_DEFAULT, w=WIDTH, h=HEIGHT, out=None):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    t = np.ascontiguousarray(t, dtype=np.float64).reshape(3)
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    if out is None:
        out = np.empty((h, w), dtype=np.float32)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().synth_render_depth(R.ctypes.data_as(dp), t.ctypes.data_as(dp), K.ctypes.data_as(dp), w, h,
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def clearance(p):
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(3)
    return lib().synth_clearance(p.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))


def render_sequence(n_frames, start=0, K=K_DEFAULT, w=WIDTH, h=HEIGHT, out=None):
    """-> (depth [n,h,w] float32, R [n,3,3], t [n,3]) along the GT path (ping-pong beyond its end)."""
    _, Rs, ts = load_trajectory()
    idx = [frame_index(start + i, len(Rs)) for i in range(n_frames)]
    if out is None:
        out = np.empty((n_frames, h, w), dtype=np.float32)
    for q, i in enumerate(idx):
        render_depth(Rs[i], ts[i], K, w, h, out[q])
    return out, Rs[idx].copy(), ts[idx].copy()


def add_sensor_noise(depth, seed=1234, dropout=0.002):
    """Noisy mode for robustness runs (fixed seed): axial Kinect-style noise, sigma_z(z) =
    0.0012 + 0.0019 (z - 0.4)^2 m (Nguyen et al. 2012), plus a small fraction of dropped (NaN) pixels.
    Deterministic for a given (frame contents, seed); accepts [h,w] or [n,h,w]; returns float32."""
    d = np.asarray(depth, np.float32)
    rng = np.random.default_rng(seed)
    sigma = (0.0012 + 0.0019 * (d.astype(np.float64) - 0.4) ** 2)
    out = (d + rng.standard_normal(d.shape) * sigma).astype(np.float32)
    if dropout > 0:
        out[rng.random(d.shape) < dropout] = np.nan
    return out


def synth_rgb(depth, R, t, K=None):
    """Procedural colour image registered to a depth frame: a 3-D checker/gradient texture evaluated at
    the world position of every pixel (so the same surface point keeps its colour across frames).
    Returns (h, w, 3) uint8; pixels with invalid depth are black."""
    K = K_DEFAULT if K is None else np.asarray(K, float).reshape(9)
    h, w = depth.shape
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    z = depth.astype(np.float64)
    cam = np.stack([(u - K[2]) * z / K[0], (v - K[5]) * z / K[4], z], axis=-1)
    world = cam @ np.asarray(R, float).T + np.asarray(t, float)
    ok = np.isfinite(z) & (z > 0)
    world = np.where(ok[..., None], world, 0.0)
    cell = np.floor(world / 0.25).astype(np.int64)
    check = ((cell[..., 0] + cell[..., 1] + cell[..., 2]) & 1).astype(np.float64)
    rgb = np.empty((h, w, 3), np.float64)
    rgb[..., 0] = 40 + 150 * check + 60 * (0.5 + 0.5 * np.sin(3.0 * world[..., 0]))
    rgb[..., 1] = 30 + 200 * (0.5 + 0.5 * np.sin(2.0 * world[..., 1] + 1.0))
    rgb[..., 2] = 255 * np.clip(world[..., 2] / 2.5, 0, 1) * (1.0 - 0.5 * check)
    rgb = np.clip(np.rint(rgb), 0, 255).astype(np.uint8)
    rgb[~ok] = 0
    return rgb
