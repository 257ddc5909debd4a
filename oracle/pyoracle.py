"""ctypes binding of the CPU oracle (oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the checker or the timed CPU baseline.  The product package
(tracking_sdf_b200) never imports this module.  Pinned against the reference itself: oracle/pyref.py binds the
reference's own translation units (compiled over oracle/shim/) and tests/test_oracle_vs_ref.py compares the two.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

c_dp = ctypes.POINTER(ctypes.c_double)
c_fp = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)


class Config(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int32), ("width", ctypes.c_float), ("height", ctypes.c_float), ("depth", ctypes.c_float),
                ("origin", ctypes.c_double * 3), ("distance_delta", ctypes.c_float), ("distance_epsilon", ctypes.c_float),
                ("gauss_newton_max_iteration", ctypes.c_int32), ("maximum_twist_diff", ctypes.c_float),
                ("v_h", ctypes.c_float), ("w_h", ctypes.c_float), ("pixel_stride", ctypes.c_int32),
                ("metric", ctypes.c_int32), ("image_width", ctypes.c_int32), ("image_height", ctypes.c_int32),
                ("use_coord_table", ctypes.c_int32), ("preprocess", ctypes.c_int32)]


class TrackStats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("stopped", ctypes.c_int32), ("n_valid", ctypes.c_int32),
                ("n_oob", ctypes.c_int32), ("singular", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("residual", ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad"}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        from tools import build
        path = build.build_oracle()
    L = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    L.orc_default_config.argtypes = [ctypes.POINTER(Config)]
    L.orc_create.argtypes = [ctypes.POINTER(Config)]; L.orc_create.restype = vp
    L.orc_destroy.argtypes = [vp]
    L.orc_set_intrinsics.argtypes = [vp, c_dp]
    L.orc_set_pose.argtypes = [vp, c_dp, c_dp]
    L.orc_get_pose.argtypes = [vp, c_dp, c_dp]
    L.orc_get_pose_inv.argtypes = [vp, c_dp, c_dp]
    L.orc_backproject.argtypes = [vp, c_fp, c_fp, c_fp]
    L.orc_preprocess.argtypes = [vp, c_fp, c_fp, c_fp]
    L.orc_k0_normals.argtypes = [vp, c_fp, c_fp]
    L.orc_fuse.argtypes = [vp, c_fp]; L.orc_fuse.restype = ctypes.c_int64
    L.orc_fuse_cloud.argtypes = [vp, c_fp, c_fp]; L.orc_fuse_cloud.restype = ctypes.c_int64
    L.orc_fuse_rgb.argtypes = [vp, c_fp, c_u8p]; L.orc_fuse_rgb.restype = ctypes.c_int64
    L.orc_enable_color.argtypes = [vp]
    L.orc_color.argtypes = [vp, ctypes.c_int]; L.orc_color.restype = c_fp
    L.orc_interpolate_color.argtypes = [vp, ctypes.c_int64, c_dp, c_fp]
    L.orc_mesh.argtypes = [vp, ctypes.c_float]; L.orc_mesh.restype = ctypes.c_int64
    L.orc_mesh_copy.argtypes = [vp, c_fp, c_dp, c_fp]
    L.orc_track.argtypes = [vp, c_fp, ctypes.POINTER(TrackStats)]
    L.orc_linearize.argtypes = [vp, c_fp, c_dp, c_dp, ctypes.POINTER(TrackStats)]
    L.orc_linearize_pixels.argtypes = [vp, c_fp, c_fp, c_fp, c_u8p]; L.orc_linearize_pixels.restype = ctypes.c_int32
    L.orc_apply_update.argtypes = [vp, c_dp, c_dp, c_dp]; L.orc_apply_update.restype = ctypes.c_int32
    L.orc_interpolate.argtypes = [vp, ctypes.c_int64, c_dp, c_fp, c_u8p]
    L.orc_exp_map.argtypes = [c_dp, c_dp, c_dp]
    L.orc_get_array_index.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]; L.orc_get_array_index.restype = ctypes.c_int64
    L.orc_get_voxel_coordinates_idx.argtypes = [vp, ctypes.c_int64, c_i32p]
    L.orc_get_voxel_coordinates.argtypes = [vp, c_dp, c_dp]
    L.orc_get_global_coordinates.argtypes = [vp, c_i32p, c_dp]
    L.orc_D.argtypes = [vp]; L.orc_D.restype = c_fp
    L.orc_W.argtypes = [vp]; L.orc_W.restype = c_fp
    L.orc_number_of_voxels.argtypes = [vp]; L.orc_number_of_voxels.restype = ctypes.c_int64
    L.orc_reset.argtypes = [vp]
    L.orc_create_circle.argtypes = [vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    L.orc_get_constants.argtypes = [vp, c_fp]
    L.orc_num_threads.restype = ctypes.c_int
    L.orc_set_num_threads.argtypes = [ctypes.c_int]
    _lib = L
    return L


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f(a):
    return a.ctypes.data_as(c_fp)


def default_config(**kw):
    c = Config()
    lib().orc_default_config(ctypes.byref(c))
    for k, v in kw.items():
        if k == "origin":
            for q in range(3):
                c.origin[q] = float(v[q])
        else:
            assert hasattr(c, k), k
            setattr(c, k, v)
    return c


class Oracle:
    """The reference's SDF + CameraTracking pair, restated (one object holds both)."""

    def __init__(self, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.L = lib()
        self.h = self.L.orc_create(ctypes.byref(self.cfg))
        self.m = self.cfg.m
        self.w, self.hgt = self.cfg.image_width, self.cfg.image_height

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # camera_tracking.cpp:22-36
    def set_intrinsics(self, K):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        self.L.orc_set_intrinsics(self.h, _d(K))

    # camera_tracking.cpp:59-65
    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        self.L.orc_set_pose(self.h, _d(R), _d(t))

    def get_pose(self):
        R = np.empty(9); t = np.empty(3)
        self.L.orc_get_pose(self.h, _d(R), _d(t))
        return R.reshape(3, 3), t

    def get_pose_inv(self):
        R = np.empty(9); t = np.empty(3)
        self.L.orc_get_pose_inv(self.h, _d(R), _d(t))
        return R.reshape(3, 3), t

    def backproject(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        cloud = np.empty((self.hgt, self.w, 3), np.float32)
        normals = np.empty((self.hgt, self.w, 3), np.float32)
        self.L.orc_backproject(self.h, _f(depth), _f(cloud), _f(normals))
        return cloud, normals

    # K0: the node's pre-processing (sdf_reconstruction.cpp:37-49), own definition after PCL's algorithm
    def preprocess(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        out = np.empty((self.hgt, self.w), np.float32)
        normals = np.empty((self.hgt, self.w, 3), np.float32)
        self.L.orc_preprocess(self.h, _f(depth), _f(out), _f(normals))
        return out, normals

    def k0_normals(self, depth_filtered):
        d = np.ascontiguousarray(depth_filtered, np.float32)
        normals = np.empty((self.hgt, self.w, 3), np.float32)
        self.L.orc_k0_normals(self.h, _f(d), _f(normals))
        return normals

    # sdf.cpp:224-305
    def fuse(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        return self.L.orc_fuse(self.h, _f(depth))

    # sdf.cpp:224-305 with the colour running mean (:294-304); rgb: (h, w, 3) uint8
    def fuse_rgb(self, depth, rgb):
        depth = np.ascontiguousarray(depth, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        assert rgb.shape[-1] == 3 and rgb.size == depth.size * 3
        return self.L.orc_fuse_rgb(self.h, _f(depth), rgb.ctypes.data_as(c_u8p))

    def enable_color(self):
        self.L.orc_enable_color(self.h)

    def color(self):
        """(Color_W, R, G, B) views, each [i, j, k] (reference layout)."""
        return tuple(self._grid(lambda h, q=q: self.L.orc_color(h, q)) for q in range(4))

    # sdf.cpp:164-217
    def interpolate_color(self, global_pts):
        pts = np.ascontiguousarray(global_pts, np.float64).reshape(-1, 3)
        out = np.empty((len(pts), 4), np.float32)
        self.L.orc_interpolate_color(self.h, len(pts), _d(pts), _f(out))
        return out

    # marching_cubes_sdf.cpp:243-287 (+ sdf.cpp:354-356, 380-385 when world / colours are asked for)
    def mesh(self, iso_level=0.0, world=False, colors=False):
        n = self.L.orc_mesh(self.h, ctypes.c_float(iso_level))
        xyz = np.empty((n, 3), np.float32)
        wd = np.empty((n, 3), np.float64) if world else None
        col = np.empty((n, 4), np.float32) if colors else None
        self.L.orc_mesh_copy(self.h, _f(xyz), _d(wd) if world else None, _f(col) if colors else None)
        return (xyz,) + ((wd,) if world else ()) + ((col,) if colors else ())

    # camera_tracking.cpp:66-245
    def track(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        st = TrackStats()
        self.L.orc_track(self.h, _f(depth), ctypes.byref(st))
        return st.as_dict()

    def linearize(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        A = np.empty(36); b = np.empty(6); st = TrackStats()
        self.L.orc_linearize(self.h, _f(depth), _d(A), _d(b), ctypes.byref(st))
        return A.reshape(6, 6), b, st.as_dict()

    def n_strided(self):
        s = self.cfg.pixel_stride
        return ((self.w + s - 1) // s) * ((self.hgt + s - 1) // s)

    def linearize_pixels(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        n = self.n_strided()
        J = np.empty((n, 6), np.float32); psi = np.empty(n, np.float32); flag = np.empty(n, np.uint8)
        r = self.L.orc_linearize_pixels(self.h, _f(depth), _f(J), _f(psi), flag.ctypes.data_as(c_u8p))
        assert r == n
        return J, psi, flag

    def apply_update(self, A, b):
        A = np.ascontiguousarray(A, np.float64).reshape(36); b = np.ascontiguousarray(b, np.float64).reshape(6)
        tw = np.empty(6)
        sing = self.L.orc_apply_update(self.h, _d(A), _d(b), _d(tw))
        return tw, sing

    # sdf.cpp:127-163
    def interpolate_distance(self, pts):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.empty(len(pts), np.float32); ok = np.empty(len(pts), np.uint8)
        self.L.orc_interpolate(self.h, len(pts), _d(pts), _f(out), ok.ctypes.data_as(c_u8p))
        return out, ok.astype(bool)

    def get_array_index(self, i, j, k):
        return self.L.orc_get_array_index(self.h, i, j, k)

    def get_voxel_coordinates_idx(self, idx):
        o = np.empty(3, np.int32)
        self.L.orc_get_voxel_coordinates_idx(self.h, idx, o.ctypes.data_as(c_i32p))
        return o

    def get_voxel_coordinates(self, g):
        g = np.ascontiguousarray(g, np.float64).reshape(3); v = np.empty(3)
        self.L.orc_get_voxel_coordinates(self.h, _d(g), _d(v))
        return v

    def get_global_coordinates(self, ijk):
        q = np.ascontiguousarray(ijk, np.int32).reshape(3); g = np.empty(3)
        self.L.orc_get_global_coordinates(self.h, q.ctypes.data_as(c_i32p), _d(g))
        return g

    def _grid(self, fn):
        n = self.L.orc_number_of_voxels(self.h)
        p = fn(self.h)
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape(self.m, self.m, self.m)   # [i (x), j (y), k (z)] — reference z-fastest

    @property
    def D(self):
        return self._grid(self.L.orc_D)

    @property
    def W(self):
        return self._grid(self.L.orc_W)

    def reset(self):
        self.L.orc_reset(self.h)

    def create_circle(self, radius, cx, cy, cz):
        self.L.orc_create_circle(self.h, radius, cx, cy, cz)

    def constants(self):
        c = np.empty(6, np.float32)
        self.L.orc_get_constants(self.h, _f(c))
        return c


def exp_map(twist):
    tw = np.ascontiguousarray(twist, np.float64).reshape(6)
    R = np.empty(9); t = np.empty(3)
    lib().orc_exp_map(_d(tw), _d(R), _d(t))
    return R.reshape(3, 3), t


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
