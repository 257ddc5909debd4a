"""ctypes binding of oracle/_ref/libtsdf_ref.so — THE REFERENCE ITSELF (its four hot-path translation units,
compiled unmodified against oracle/shim/, see oracle/Makefile target `ref` and oracle/ref_bridge.cpp).

TEST INFRASTRUCTURE ONLY, same rule as pyoracle: tests/, the golden generator and bench.py's --impl reference /
cpu_baseline legs use it; the product package never imports it.  It exists to PIN the oracle restatement:
tests/test_oracle_vs_ref.py compares the two function by function.

/root/reference does not exist on the GPU box: the library is built in the build container (build()) and
travels with the snapshot; `available()` says whether it is there.  The class mirrors pyoracle.Oracle.  The
reference starts from point clouds + normals (they come from ROS/PCL upstream of it), so depth images are
back-projected with the oracle's K1 definition first (orc_backproject — the one definition oracle and GPU
share); everything downstream is the reference's own code.
"""
import ctypes
import os

import numpy as np

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
PATH = os.path.join(ROOT, "oracle", "_ref", "libtsdf_ref.so")

c_dp, c_fp, c_u8p, c_i32p = po.c_dp, po.c_fp, po.c_u8p, po.c_i32p
_lib = None


def sources_present():
    return os.path.exists(os.path.join(REF_SRC, "src", "sdf.cpp"))


def build(force=False, verbose=False):
    """make -C oracle ref  (only possible where /root/reference exists)."""
    import subprocess
    if not sources_present():
        return PATH if os.path.exists(PATH) else None
    cmd = ["make", "-C", os.path.join(ROOT, "oracle"), "ref"] + (["-B"] if force else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building oracle/_ref failed")
    return PATH


def available():
    return os.path.exists(PATH) or sources_present()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if sources_present():
        build()
    if not os.path.exists(PATH):
        raise RuntimeError("oracle/_ref/libtsdf_ref.so is absent and /root/reference is not here to build it")
    L = ctypes.CDLL(PATH)
    vp = ctypes.c_void_p
    L.ref_create.argtypes = [ctypes.POINTER(po.Config)]; L.ref_create.restype = vp
    L.ref_destroy.argtypes = [vp]
    L.ref_set_intrinsics.argtypes = [vp, c_dp]
    L.ref_set_pose.argtypes = [vp, c_dp, c_dp]
    L.ref_get_pose.argtypes = [vp, c_dp, c_dp]
    L.ref_get_pose_inv.argtypes = [vp, c_dp, c_dp]
    L.ref_set_gn.argtypes = [vp, ctypes.c_int, ctypes.c_float]
    L.ref_fuse_cloud.argtypes = [vp, c_fp, c_fp, c_u8p, ctypes.c_int]; L.ref_fuse_cloud.restype = ctypes.c_int64
    L.ref_track_cloud.argtypes = [vp, c_fp, ctypes.POINTER(po.TrackStats)]
    L.ref_linearize_cloud.argtypes = [vp, c_fp, c_dp, c_dp]
    L.ref_apply_update.argtypes = [vp, c_dp, c_dp, c_dp]
    L.ref_linearize_pixels.argtypes = [vp, c_fp, c_fp, c_fp, c_u8p]; L.ref_linearize_pixels.restype = ctypes.c_int32
    L.ref_interpolate.argtypes = [vp, ctypes.c_int64, c_dp, c_fp, c_u8p]
    L.ref_exp_map.argtypes = [c_dp, c_dp, c_dp]
    L.ref_get_array_index.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]; L.ref_get_array_index.restype = ctypes.c_int64
    L.ref_get_voxel_coordinates_idx.argtypes = [vp, ctypes.c_int64, c_i32p]
    L.ref_get_voxel_coordinates.argtypes = [vp, c_dp, c_dp]
    L.ref_get_global_coordinates.argtypes = [vp, c_i32p, c_dp]
    L.ref_D.argtypes = [vp]; L.ref_D.restype = c_fp
    L.ref_W.argtypes = [vp]; L.ref_W.restype = c_fp
    L.ref_color.argtypes = [vp, ctypes.c_int]; L.ref_color.restype = c_fp
    L.ref_number_of_voxels.argtypes = [vp]; L.ref_number_of_voxels.restype = ctypes.c_int64
    L.ref_create_circle.argtypes = [vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    L.ref_interpolate_color.argtypes = [vp, ctypes.c_int64, c_dp, c_fp]
    L.ref_mesh.argtypes = [vp, ctypes.c_float]; L.ref_mesh.restype = ctypes.c_int64
    L.ref_mesh_copy.argtypes = [vp, c_fp]
    L.ref_visualize.argtypes = [vp]; L.ref_visualize.restype = ctypes.c_int64
    L.ref_marker_copy.argtypes = [vp, c_dp, c_fp]
    L.ref_get_constants.argtypes = [vp, c_fp]
    L.ref_timers.argtypes = [vp, c_dp, ctypes.c_int]
    L.ref_num_threads.restype = ctypes.c_int
    L.ref_set_num_threads.argtypes = [ctypes.c_int]
    _lib = L
    return L


_d, _f = po._d, po._f


class Reference:
    """The reference's own SDF + CameraTracking objects (sdf_reconstruction.cpp:83-88)."""

    def __init__(self, cfg=None, **kw):
        kw.pop("use_coord_table", None)            # the reference always builds its global_coords table
        self.cfg = cfg if cfg is not None else po.default_config(**kw)
        assert self.cfg.metric == 0, "the reference fuses point-to-plane only (sdf.cpp:267 is commented out)"
        assert self.cfg.pixel_stride == 3, "the reference's pixel stride is hard-coded (camera_tracking.cpp:162-163)"
        self.L = lib()
        self.h = self.L.ref_create(ctypes.byref(self.cfg))
        self.m = self.cfg.m
        self.w, self.hgt = self.cfg.image_width, self.cfg.image_height
        # K1 (back-projection + normals) is upstream of the reference: the oracle's shared definition
        self._k1 = po.Oracle(po.default_config(m=4, use_coord_table=0, image_width=self.w, image_height=self.hgt))

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None
            self._k1.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_intrinsics(self, K):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        self.L.ref_set_intrinsics(self.h, _d(K))
        self._k1.set_intrinsics(K)

    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        self.L.ref_set_pose(self.h, _d(R), _d(t))

    def get_pose(self):
        R = np.empty(9); t = np.empty(3)
        self.L.ref_get_pose(self.h, _d(R), _d(t))
        return R.reshape(3, 3), t

    def get_pose_inv(self):
        R = np.empty(9); t = np.empty(3)
        self.L.ref_get_pose_inv(self.h, _d(R), _d(t))
        return R.reshape(3, 3), t

    def set_gn(self, max_iter, max_twist_diff):
        self.L.ref_set_gn(self.h, int(max_iter), ctypes.c_float(max_twist_diff))

    def backproject(self, depth):
        return self._k1.backproject(depth)

    # SDF::update, sdf.cpp:224-315
    def fuse_cloud(self, cloud, normals, rgb=None, count=True):
        cloud = np.ascontiguousarray(cloud, np.float32); normals = np.ascontiguousarray(normals, np.float32)
        rp = None
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, np.uint8)
            rp = rgb.ctypes.data_as(c_u8p)
        return self.L.ref_fuse_cloud(self.h, _f(cloud), _f(normals), rp, 1 if count else 0)

    def fuse(self, depth, count=True):
        cloud, normals = self.backproject(depth)
        return self.fuse_cloud(cloud, normals, None, count)

    def fuse_rgb(self, depth, rgb, count=True):
        cloud, normals = self.backproject(depth)
        return self.fuse_cloud(cloud, normals, rgb, count)

    # CameraTracking::estimate_new_position, camera_tracking.cpp:66-245
    def track_cloud(self, cloud):
        cloud = np.ascontiguousarray(cloud, np.float32)
        st = po.TrackStats()
        self.L.ref_track_cloud(self.h, _f(cloud), ctypes.byref(st))
        return st.as_dict()

    def track(self, depth):
        return self.track_cloud(self.backproject(depth)[0])

    def linearize(self, depth):
        cloud = self.backproject(depth)[0]
        A = np.empty(36); b = np.empty(6)
        self.L.ref_linearize_cloud(self.h, _f(cloud), _d(A), _d(b))
        return A.reshape(6, 6), b

    def apply_update(self, A, b):
        A = np.ascontiguousarray(A, np.float64).reshape(36); b = np.ascontiguousarray(b, np.float64).reshape(6)
        tw = np.empty(6)
        self.L.ref_apply_update(self.h, _d(A), _d(b), _d(tw))
        return tw

    def n_strided(self):
        return ((self.w + 2) // 3) * ((self.hgt + 2) // 3)

    def linearize_pixels(self, depth):
        cloud = self.backproject(depth)[0]
        n = self.n_strided()
        J = np.empty((n, 6), np.float32); psi = np.empty(n, np.float32); flag = np.empty(n, np.uint8)
        r = self.L.ref_linearize_pixels(self.h, _f(cloud), _f(J), _f(psi), flag.ctypes.data_as(c_u8p))
        assert r == n
        return J, psi, flag

    def interpolate_distance(self, pts):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.empty(len(pts), np.float32); ok = np.empty(len(pts), np.uint8)
        self.L.ref_interpolate(self.h, len(pts), _d(pts), _f(out), ok.ctypes.data_as(c_u8p))
        return out, ok.astype(bool)

    def interpolate_color(self, global_pts):
        pts = np.ascontiguousarray(global_pts, np.float64).reshape(-1, 3)
        out = np.empty((len(pts), 4), np.float32)
        self.L.ref_interpolate_color(self.h, len(pts), _d(pts), _f(out))
        return out

    def mesh(self, iso_level=0.0):
        n = self.L.ref_mesh(self.h, ctypes.c_float(iso_level))
        xyz = np.empty((n, 3), np.float32)
        self.L.ref_mesh_copy(self.h, _f(xyz))
        return xyz

    def visualize(self):
        """One pass of SDF::visualize's loop (sdf.cpp:324-389): marker points (world, double) + vertex colours."""
        n = self.L.ref_visualize(self.h)
        world = np.empty((n, 3), np.float64); rgba = np.empty((n, 4), np.float32)
        self.L.ref_marker_copy(self.h, _d(world), _f(rgba))
        return world, rgba

    def get_array_index(self, i, j, k):
        return self.L.ref_get_array_index(self.h, i, j, k)

    def get_voxel_coordinates_idx(self, idx):
        o = np.empty(3, np.int32)
        self.L.ref_get_voxel_coordinates_idx(self.h, idx, o.ctypes.data_as(c_i32p))
        return o

    def get_voxel_coordinates(self, g):
        g = np.ascontiguousarray(g, np.float64).reshape(3); v = np.empty(3)
        self.L.ref_get_voxel_coordinates(self.h, _d(g), _d(v))
        return v

    def get_global_coordinates(self, ijk):
        q = np.ascontiguousarray(ijk, np.int32).reshape(3); g = np.empty(3)
        self.L.ref_get_global_coordinates(self.h, q.ctypes.data_as(c_i32p), _d(g))
        return g

    def _grid(self, p):
        n = self.L.ref_number_of_voxels(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.m, self.m, self.m)

    @property
    def D(self):
        return self._grid(self.L.ref_D(self.h))

    @property
    def W(self):
        return self._grid(self.L.ref_W(self.h))

    def color(self):
        return tuple(self._grid(self.L.ref_color(self.h, q)) for q in range(4))

    def create_circle(self, radius, cx, cy, cz):
        self.L.ref_create_circle(self.h, radius, cx, cy, cz)

    def constants(self):
        c = np.empty(6, np.float32)
        self.L.ref_get_constants(self.h, _f(c))
        return c

    def timers(self, reset=False):
        """Seconds spent inside (estimate_new_position, update) — the spans the reference prints itself."""
        t = np.empty(2)
        self.L.ref_timers(self.h, _d(t), 1 if reset else 0)
        return float(t[0]), float(t[1])


def exp_map(twist):
    tw = np.ascontiguousarray(twist, np.float64).reshape(6)
    R = np.empty(9); t = np.empty(3)
    lib().ref_exp_map(_d(tw), _d(R), _d(t))
    return R.reshape(3, 3), t


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))
