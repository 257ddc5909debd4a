"""ctypes binding of include/tsdf_b200.h (one Python method per C entry point)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

POINT_TO_PLANE, POINT_TO_POINT = 0, 1
HOST, DEVICE = 0, 1
LAYOUT_REFERENCE, LAYOUT_XFASTEST = 0, 1
IPC_HANDLE_BYTES = 64

c_dp = ctypes.POINTER(ctypes.c_double)
c_fp = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)

STATUS_NAMES = {0: "OK", 1: "BAD_ARG", 2: "NO_INTRINSICS", 3: "CUDA", 4: "TRACKING_LOST", 5: "HALO", 6: "NOMEM", 7: "PEER"}


class Config(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int32), ("width", ctypes.c_float), ("height", ctypes.c_float), ("depth", ctypes.c_float),
                ("origin", ctypes.c_double * 3), ("distance_delta", ctypes.c_float), ("distance_epsilon", ctypes.c_float),
                ("gauss_newton_max_iteration", ctypes.c_int32), ("maximum_twist_diff", ctypes.c_float),
                ("v_h", ctypes.c_float), ("w_h", ctypes.c_float), ("pixel_stride", ctypes.c_int32),
                ("metric", ctypes.c_int32), ("image_width", ctypes.c_int32), ("image_height", ctypes.c_int32),
                ("device", ctypes.c_int32), ("n_shards", ctypes.c_int32), ("shard_rank", ctypes.c_int32),
                ("halo", ctypes.c_int32), ("slab_k_begin", ctypes.c_int32), ("slab_k_end", ctypes.c_int32),
                ("preprocess", ctypes.c_int32), ("reserved", ctypes.c_int32 * 1)]


class TrackStats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int32), ("stopped", ctypes.c_int32), ("n_valid", ctypes.c_int32),
                ("n_oob", ctypes.c_int32), ("singular", ctypes.c_int32), ("halo_miss", ctypes.c_int32),
                ("residual", ctypes.c_double), ("A", ctypes.c_double * 36), ("b", ctypes.c_double * 6),
                ("twist", ctypes.c_double * 6)]

    def as_dict(self):
        return {"iterations": self.iterations, "stopped": self.stopped, "n_valid": self.n_valid, "n_oob": self.n_oob,
                "singular": self.singular, "halo_miss": self.halo_miss, "residual": self.residual,
                "A": np.array(self.A[:]).reshape(6, 6), "b": np.array(self.b[:]), "twist": np.array(self.twist[:])}


class TsdfError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("tsdf_b200: %s (%s)" % (msg, STATUS_NAMES.get(status, status)))
        self.status = status


# every symbol include/tsdf_b200.h declares: (name, restype, argtypes)
_VP = ctypes.c_void_p
_VPP = ctypes.POINTER(ctypes.c_void_p)
_CFGP = ctypes.POINTER(Config)
_STP = ctypes.POINTER(TrackStats)
_I32, _I64 = ctypes.c_int32, ctypes.c_int64
PROTOTYPES = [
    ("tsdf_abi_version", _I32, []),
    ("tsdf_last_error", ctypes.c_char_p, []),
    ("tsdf_device_count", _I32, []),
    ("tsdf_default_config", None, [_CFGP]),
    ("tsdf_create", _I32, [_CFGP, _VPP]),
    ("tsdf_destroy", _I32, [_VP]),
    ("tsdf_reset", _I32, [_VP]),
    ("tsdf_get_config", _I32, [_VP, _CFGP]),
    ("tsdf_set_intrinsics", _I32, [_VP, c_dp]),
    ("tsdf_set_pose", _I32, [_VP, c_dp, c_dp]),
    ("tsdf_get_pose", _I32, [_VP, c_dp, c_dp]),
    ("tsdf_get_pose_inv", _I32, [_VP, c_dp, c_dp]),
    ("tsdf_track", _I32, [_VP, _VP, _I32, c_dp, c_dp, _STP]),
    ("tsdf_fuse", _I32, [_VP, _VP, _I32, c_dp, c_dp, c_i64p]),
    ("tsdf_track_and_fuse", _I32, [_VP, _VP, _I32, c_dp, c_dp, _STP, c_i64p]),
    ("tsdf_enable_color", _I32, [_VP]),
    ("tsdf_fuse_rgb", _I32, [_VP, _VP, _VP, _I32, c_dp, c_dp, c_i64p]),
    ("tsdf_track_and_fuse_rgb", _I32, [_VP, _VP, _VP, _I32, c_dp, c_dp, _STP, c_i64p]),
    ("tsdf_interpolate_color", _I32, [_VP, _I64, c_dp, c_fp]),
    ("tsdf_download_color", _I32, [_VP, c_fp, c_fp, c_fp, c_fp, _I32]),
    ("tsdf_balanced_slabs", _I32, [_I32, _I32, c_dp, _I32, _I32, ctypes.POINTER(ctypes.c_int32)]),
    ("tsdf_balanced_slabs2", _I32, [_I32, _I32, c_dp, c_dp, _I32, _I32, ctypes.POINTER(ctypes.c_int32)]),
    ("tsdf_mesh_extract", _I32, [_VP, ctypes.c_float, c_i64p]),
    ("tsdf_mesh_download", _I32, [_VP, c_fp, c_dp, c_fp]),
    ("tsdf_enqueue_frame", _I32, [_VP, _VP, _I32, _I32]),
    ("tsdf_submit_frame", _I32, [_VP, _VP, _I32, _I32]),
    ("tsdf_sync", _I32, [_VP]),
    ("tsdf_pose_ring_capacity", _I32, []),
    ("tsdf_read_pose_ring", _I32, [_VP, _I32, c_dp, c_dp, _STP]),
    ("tsdf_linearize", _I32, [_VP, _VP, _I32, c_dp, c_dp, _STP]),
    ("tsdf_num_strided_pixels", _I32, [_VP]),
    ("tsdf_linearize_pixels", _I32, [_VP, _VP, _I32, c_fp, c_fp, c_u8p]),
    ("tsdf_backproject", _I32, [_VP, _VP, _I32, c_fp, c_fp]),
    ("tsdf_interpolate_distance", _I32, [_VP, _I64, c_dp, c_fp, c_u8p]),
    ("tsdf_number_of_voxels", _I64, [_VP]),
    ("tsdf_stored_range", _I32, [_VP, c_i32p, c_i32p, c_i32p, c_i32p]),
    ("tsdf_download", _I32, [_VP, c_fp, c_fp, _I32]),
    ("tsdf_upload", _I32, [_VP, c_fp, c_fp, _I32]),
    ("tsdf_device_grid", _I32, [_VP, _VPP, c_i64p]),
    ("tsdf_get_array_index", _I64, [_VP, _I32, _I32, _I32]),
    ("tsdf_get_voxel_coordinates_idx", None, [_VP, _I64, c_i32p]),
    ("tsdf_get_voxel_coordinates", None, [_VP, c_dp, c_dp]),
    ("tsdf_get_global_coordinates", None, [_VP, c_i32p, c_dp]),
    ("tsdf_exp_map", _I32, [_VP, c_dp, c_dp, c_dp]),
    ("tsdf_dev_alloc", _I32, [_VP, _I64, _VPP]),
    ("tsdf_dev_free", _I32, [_VP, _VP]),
    ("tsdf_dev_upload", _I32, [_VP, _VP, _VP, _I64]),
    ("tsdf_host_alloc_pinned", _I32, [_I64, _VPP]),
    ("tsdf_host_free_pinned", _I32, [_VP]),
    ("tsdf_last_stage_ms", _I32, [_VP, c_fp]),
    ("tsdf_event_timer_begin", _I32, [_VP]),
    ("tsdf_event_timer_end", _I32, [_VP, c_fp]),
    ("tsdf_kernel_launch_count", _I64, [_VP]),
    ("tsdf_flush_l2", _I32, [_VP]),
    ("tsdf_stage_timing_begin", _I32, [_VP, _I32]),
    ("tsdf_stage_timing_end", _I32, [_VP, c_i32p, c_fp]),
    ("tsdf_total_updates", _I32, [_VP, _I32, c_i64p]),
    ("tsdf_debug_phase_times", _I32, [_VP, _VP, _I32, c_i64p]),
    ("tsdf_debug_stream_rmw", _I32, [_VP, _I32, c_fp]),
    ("tsdf_debug_fuse_check", _I32, [_VP, _VP, _I32, c_i64p]),
    ("tsdf_preprocess", _I32, [_VP, ctypes.c_void_p, _I32, c_fp, c_fp]),
    ("tsdf_debug_check_rcp", _I32, [_VP, ctypes.c_float, ctypes.c_float, c_i64p]),
    ("tsdf_debug_check_weight_exp", _I32, [_VP, ctypes.c_float, ctypes.c_float, c_i64p, c_fp, c_fp, _I32]),
    ("tsdf_slab_plan", _I32, [_CFGP, c_i32p]),
    ("tsdf_shard_ipc_export", _I32, [_VP, c_u8p]),
    ("tsdf_shard_ipc_attach", _I32, [_VP, _I32, c_u8p]),
    ("tsdf_shard_attach_local", _I32, [_VPP, _I32]),
    ("tsdf_group_set_intrinsics", _I32, [_VPP, _I32, c_dp]),
    ("tsdf_group_set_pose", _I32, [_VPP, _I32, c_dp, c_dp]),
    ("tsdf_group_linearize", _I32, [_VPP, _I32, _VP, _I32, c_dp, c_dp, _STP]),
    ("tsdf_group_frame", _I32, [_VPP, _I32, _VP, _I32, _I32, _I32, c_dp, c_dp, _STP, c_i64p]),
]

_lib = None


def library_path():
    return os.path.join(_HERE, "_lib", "libtsdf_b200.so")


def load_library():
    """Load libtsdf_b200.so.  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TSDF_B200_LIB", library_path())      # override: kernel tuning experiments only
    if not os.path.exists(path):
        raise TsdfError(3, "libtsdf_b200.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "tracking_sdf_b200 has no CPU fallback")
    L = ctypes.CDLL(path)
    for name, res, args in PROTOTYPES:
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def default_config(**kw):
    c = Config()
    load_library().tsdf_default_config(ctypes.byref(c))
    for k, v in kw.items():
        if k == "origin":
            for q in range(3):
                c.origin[q] = float(v[q])
        else:
            assert hasattr(c, k), k
            setattr(c, k, v)
    return c


def slab_plan(cfg):
    """-> dict(own=(k0,k1), stored=(k0,k1), halo=h): pure host computation, no GPU needed."""
    L = load_library()
    out = (ctypes.c_int32 * 5)()
    st = L.tsdf_slab_plan(ctypes.byref(cfg), out)
    if st != 0:
        raise TsdfError(st, L.tsdf_last_error().decode())
    return {"own": (out[0], out[1]), "stored": (out[2], out[3]), "halo": out[4]}


def balanced_slabs(weights, n_shards, min_layers=8, halo=0, weights_own=None):
    """Cuts [b0=0, b1, ..., bn=m] of the z partition that minimises the largest per-slab cost, halo layers
    included (host only).  weights_own: a second profile that counts for a layer's owner only (tracked pixels)."""
    L = load_library()
    w = np.ascontiguousarray(weights, np.float64)
    out = (ctypes.c_int32 * (n_shards + 1))()
    if weights_own is not None:
        wo = np.ascontiguousarray(weights_own, np.float64)
        assert len(wo) == len(w)
        st = L.tsdf_balanced_slabs2(len(w), n_shards, w.ctypes.data_as(c_dp), wo.ctypes.data_as(c_dp), min_layers, halo, out)
    else:
        st = L.tsdf_balanced_slabs(len(w), n_shards, w.ctypes.data_as(c_dp), min_layers, halo, out)
    if st != 0:
        raise TsdfError(st, L.tsdf_last_error().decode())
    return [int(v) for v in out]


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f(a):
    return a.ctypes.data_as(c_fp)


def _depth_arg(depth):
    """numpy array (host) or int (device pointer) -> (void*, mem, keepalive)"""
    if isinstance(depth, (int, np.integer)):
        return ctypes.c_void_p(int(depth)), DEVICE, None
    a = np.ascontiguousarray(depth, np.float32)
    return ctypes.c_void_p(a.ctypes.data), HOST, a


class Tsdf:
    """One handle = the reference's SDF + CameraTracking pair on one GPU (or one z-slab of it)."""

    def __init__(self, cfg=None, **kw):
        self.L = load_library()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        h = ctypes.c_void_p()
        self._ck(self.L.tsdf_create(ctypes.byref(self.cfg), ctypes.byref(h)))
        self.h = h
        self.L.tsdf_get_config(self.h, ctypes.byref(self.cfg))
        self.m = self.cfg.m
        self.w, self.hgt = self.cfg.image_width, self.cfg.image_height

    def _ck(self, st):
        if st != 0:
            raise TsdfError(st, self.L.tsdf_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.tsdf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._ck(self.L.tsdf_reset(self.h))

    def set_intrinsics(self, K):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        self._ck(self.L.tsdf_set_intrinsics(self.h, _d(K)))

    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        self._ck(self.L.tsdf_set_pose(self.h, _d(R), _d(t)))

    def get_pose(self):
        R = np.empty(9); t = np.empty(3)
        self._ck(self.L.tsdf_get_pose(self.h, _d(R), _d(t)))
        return R.reshape(3, 3), t

    def get_pose_inv(self):
        R = np.empty(9); t = np.empty(3)
        self._ck(self.L.tsdf_get_pose_inv(self.h, _d(R), _d(t)))
        return R.reshape(3, 3), t

    def track(self, depth):
        p, mem, keep = _depth_arg(depth)
        R = np.empty(9); t = np.empty(3); st = TrackStats()
        self._ck(self.L.tsdf_track(self.h, p, mem, _d(R), _d(t), ctypes.byref(st)))
        return R.reshape(3, 3), t, st.as_dict()

    def fuse(self, depth, R=None, t=None):
        p, mem, keep = _depth_arg(depth)
        n = ctypes.c_int64()
        if R is None:
            self._ck(self.L.tsdf_fuse(self.h, p, mem, None, None, ctypes.byref(n)))
        else:
            R = np.ascontiguousarray(R, np.float64).reshape(9)
            t = np.ascontiguousarray(t, np.float64).reshape(3)
            self._ck(self.L.tsdf_fuse(self.h, p, mem, _d(R), _d(t), ctypes.byref(n)))
        return n.value

    def track_and_fuse(self, depth):
        p, mem, keep = _depth_arg(depth)
        R = np.empty(9); t = np.empty(3); st = TrackStats(); n = ctypes.c_int64()
        self._ck(self.L.tsdf_track_and_fuse(self.h, p, mem, _d(R), _d(t), ctypes.byref(st), ctypes.byref(n)))
        return R.reshape(3, 3), t, st.as_dict(), n.value

    # ---- colour (sdf.cpp:294-304, 164-217); rgb: (h, w, 3) uint8 host array
    def enable_color(self):
        self._ck(self.L.tsdf_enable_color(self.h))

    @staticmethod
    def _rgb_arg(rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        return ctypes.c_void_p(rgb.ctypes.data), rgb

    def fuse_rgb(self, depth, rgb, R=None, t=None):
        p, mem, keep = _depth_arg(depth)
        assert mem == HOST
        q, keep2 = self._rgb_arg(rgb)
        n = ctypes.c_int64()
        if R is None:
            self._ck(self.L.tsdf_fuse_rgb(self.h, p, q, mem, None, None, ctypes.byref(n)))
        else:
            R = np.ascontiguousarray(R, np.float64).reshape(9)
            t = np.ascontiguousarray(t, np.float64).reshape(3)
            self._ck(self.L.tsdf_fuse_rgb(self.h, p, q, mem, _d(R), _d(t), ctypes.byref(n)))
        return n.value

    def track_and_fuse_rgb(self, depth, rgb):
        p, mem, keep = _depth_arg(depth)
        assert mem == HOST
        q, keep2 = self._rgb_arg(rgb)
        R = np.empty(9); t = np.empty(3); st = TrackStats(); n = ctypes.c_int64()
        self._ck(self.L.tsdf_track_and_fuse_rgb(self.h, p, q, mem, _d(R), _d(t), ctypes.byref(st), ctypes.byref(n)))
        return R.reshape(3, 3), t, st.as_dict(), n.value

    def interpolate_color(self, global_pts):
        pts = np.ascontiguousarray(global_pts, np.float64).reshape(-1, 3)
        out = np.empty((len(pts), 4), np.float32)
        self._ck(self.L.tsdf_interpolate_color(self.h, len(pts), _d(pts), _f(out)))
        return out

    def download_color(self, layout=LAYOUT_REFERENCE):
        """-> Color_W, R, G, B, shaped like download()."""
        ks0, ks1, _, _ = self.stored_range()
        nk = ks1 - ks0
        shape = (self.m, self.m, nk) if layout == LAYOUT_REFERENCE else (nk, self.m, self.m)
        a = [np.empty(shape, np.float32) for _ in range(4)]
        self._ck(self.L.tsdf_download_color(self.h, _f(a[0]), _f(a[1]), _f(a[2]), _f(a[3]), layout))
        return tuple(a)

    # ---- mesh (marching_cubes_sdf.cpp:243-287; sdf.cpp:354-356, 380-385)
    def mesh_extract(self, iso_level=0.0):
        """Run the device mesher; the mesh stays on the device.  -> number of vertices."""
        n = ctypes.c_int64()
        self._ck(self.L.tsdf_mesh_extract(self.h, ctypes.c_float(iso_level), ctypes.byref(n)))
        return n.value

    def mesh(self, iso_level=0.0, world=False, colors=False):
        """-> (xyz [n,3] float32[, world [n,3] float64][, rgba [n,4] float32]); 3 vertices per triangle."""
        n = self.mesh_extract(iso_level)
        xyz = np.empty((n, 3), np.float32)
        wd = np.empty((n, 3), np.float64) if world else None
        col = np.empty((n, 4), np.float32) if colors else None
        self._ck(self.L.tsdf_mesh_download(self.h, _f(xyz), _d(wd) if world else None, _f(col) if colors else None))
        return (xyz,) + ((wd,) if world else ()) + ((col,) if colors else ())

    def enqueue_frame(self, depth_dev, track, slot):
        self._ck(self.L.tsdf_enqueue_frame(self.h, ctypes.c_void_p(int(depth_dev)), int(track), int(slot)))

    def submit_frame(self, depth_host, track, slot):
        a = depth_host
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        self._ck(self.L.tsdf_submit_frame(self.h, ctypes.c_void_p(a.ctypes.data), int(track), int(slot)))

    def sync(self):
        self._ck(self.L.tsdf_sync(self.h))

    def pose_ring_capacity(self):
        return self.L.tsdf_pose_ring_capacity()

    def read_pose_ring(self, slot):
        R = np.empty(9); t = np.empty(3); st = TrackStats()
        self._ck(self.L.tsdf_read_pose_ring(self.h, slot, _d(R), _d(t), ctypes.byref(st)))
        return R.reshape(3, 3), t, st.as_dict()

    def linearize(self, depth):
        p, mem, keep = _depth_arg(depth)
        A = np.empty(36); b = np.empty(6); st = TrackStats()
        self._ck(self.L.tsdf_linearize(self.h, p, mem, _d(A), _d(b), ctypes.byref(st)))
        return A.reshape(6, 6), b, st.as_dict()

    def num_strided_pixels(self):
        return self.L.tsdf_num_strided_pixels(self.h)

    def linearize_pixels(self, depth):
        p, mem, keep = _depth_arg(depth)
        n = self.num_strided_pixels()
        J = np.empty((n, 6), np.float32); psi = np.empty(n, np.float32); flag = np.empty(n, np.uint8)
        self._ck(self.L.tsdf_linearize_pixels(self.h, p, mem, _f(J), _f(psi), flag.ctypes.data_as(c_u8p)))
        return J, psi, flag

    def backproject(self, depth, normals=True):
        p, mem, keep = _depth_arg(depth)
        cloud = np.empty((self.hgt, self.w, 3), np.float32)
        nrm = np.empty((self.hgt, self.w, 3), np.float32) if normals else None
        self._ck(self.L.tsdf_backproject(self.h, p, mem, _f(cloud), _f(nrm) if normals else None))
        return cloud, nrm

    def interpolate_distance(self, pts):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.empty(len(pts), np.float32); ok = np.empty(len(pts), np.uint8)
        self._ck(self.L.tsdf_interpolate_distance(self.h, len(pts), _d(pts), _f(out), ok.ctypes.data_as(c_u8p)))
        return out, ok.astype(bool)

    def number_of_voxels(self):
        return self.L.tsdf_number_of_voxels(self.h)

    def stored_range(self):
        v = [ctypes.c_int32() for _ in range(4)]
        self._ck(self.L.tsdf_stored_range(self.h, *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def download(self, layout=LAYOUT_REFERENCE):
        """-> D, W.  LAYOUT_REFERENCE: arrays indexed [i, j, k - k_begin]; XFASTEST: [k - k_begin, j, i]."""
        ks0, ks1, _, _ = self.stored_range()
        nk = ks1 - ks0
        shape = (self.m, self.m, nk) if layout == LAYOUT_REFERENCE else (nk, self.m, self.m)
        D = np.empty(shape, np.float32); W = np.empty(shape, np.float32)
        self._ck(self.L.tsdf_download(self.h, _f(D), _f(W), layout))
        return D, W

    def upload(self, D, W, layout=LAYOUT_REFERENCE):
        D = np.ascontiguousarray(D, np.float32); W = np.ascontiguousarray(W, np.float32)
        self._ck(self.L.tsdf_upload(self.h, _f(D), _f(W), layout))

    def device_grid(self):
        p = ctypes.c_void_p(); n = ctypes.c_int64()
        self._ck(self.L.tsdf_device_grid(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def get_array_index(self, i, j, k):
        return self.L.tsdf_get_array_index(self.h, i, j, k)

    def get_voxel_coordinates_idx(self, idx):
        o = np.empty(3, np.int32)
        self.L.tsdf_get_voxel_coordinates_idx(self.h, idx, o.ctypes.data_as(c_i32p))
        return o

    def get_voxel_coordinates(self, g):
        g = np.ascontiguousarray(g, np.float64).reshape(3); v = np.empty(3)
        self.L.tsdf_get_voxel_coordinates(self.h, _d(g), _d(v))
        return v

    def get_global_coordinates(self, ijk):
        q = np.ascontiguousarray(ijk, np.int32).reshape(3); g = np.empty(3)
        self.L.tsdf_get_global_coordinates(self.h, q.ctypes.data_as(c_i32p), _d(g))
        return g

    def exp_map(self, twist):
        tw = np.ascontiguousarray(twist, np.float64).reshape(6); R = np.empty(9); t = np.empty(3)
        self._ck(self.L.tsdf_exp_map(self.h, _d(tw), _d(R), _d(t)))
        return R.reshape(3, 3), t

    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(self.L.tsdf_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, ptr):
        self._ck(self.L.tsdf_dev_free(self.h, ctypes.c_void_p(ptr)))

    def dev_upload(self, dev_ptr, host_array):
        a = np.ascontiguousarray(host_array)
        self._ck(self.L.tsdf_dev_upload(self.h, ctypes.c_void_p(dev_ptr), ctypes.c_void_p(a.ctypes.data), a.nbytes))

    def last_stage_ms(self):
        o = np.empty(3, np.float32)
        self._ck(self.L.tsdf_last_stage_ms(self.h, _f(o)))
        return o

    def timer_begin(self):
        self._ck(self.L.tsdf_event_timer_begin(self.h))

    def timer_end(self):
        ms = ctypes.c_float()
        self._ck(self.L.tsdf_event_timer_end(self.h, ctypes.byref(ms)))
        return ms.value

    def kernel_launch_count(self):
        return self.L.tsdf_kernel_launch_count(self.h)

    def stage_timing_begin(self, n_frames):
        self._stage_n = n_frames
        self._ck(self.L.tsdf_stage_timing_begin(self.h, n_frames))

    def stage_timing_end(self):
        ms = np.empty((self._stage_n, 3), np.float32); n = ctypes.c_int32()
        self._ck(self.L.tsdf_stage_timing_end(self.h, ctypes.byref(n), _f(ms)))
        return ms[:n.value]

    def total_updates(self, reset=False):
        v = ctypes.c_int64()
        self._ck(self.L.tsdf_total_updates(self.h, int(reset), ctypes.byref(v)))
        return v.value

    def debug_stream_rmw(self, reps=10):
        ms = ctypes.c_float()
        self._ck(self.L.tsdf_debug_stream_rmw(self.h, reps, ctypes.byref(ms)))
        return ms.value

    def debug_fuse_check(self, depth):
        p, mem, keep = _depth_arg(depth)
        out = (ctypes.c_int64 * 3)()
        self._ck(self.L.tsdf_debug_fuse_check(self.h, p, mem, out))
        return {"fast": int(out[0]), "wrong": int(out[1]), "items": int(out[2])}

    def debug_check_rcp(self, x_lo, x_hi):
        n = ctypes.c_int64()
        self._ck(self.L.tsdf_debug_check_rcp(self.h, x_lo, x_hi, ctypes.byref(n)))
        return n.value

    def preprocess(self, depth):
        """K0 alone: (filtered depth [h,w], normals [h,w,3]) — see tsdf_preprocess."""
        p, mem, keep = _depth_arg(depth)
        zf = np.empty((self.hgt, self.w), np.float32); n = np.empty((self.hgt, self.w, 3), np.float32)
        self._ck(self.L.tsdf_preprocess(self.h, p, mem, _f(zf), _f(n)))
        return zf, n

    def debug_check_weight_exp(self, e_lo, e_hi, cap=65536):
        """-> (n_ambiguous, e[n], w_device[n]) — see tsdf_debug_check_weight_exp."""
        n = ctypes.c_int64()
        e = np.empty(cap, np.float32); w = np.empty(cap, np.float32)
        self._ck(self.L.tsdf_debug_check_weight_exp(self.h, e_lo, e_hi, ctypes.byref(n), _f(e), _f(w), cap))
        k = min(n.value, cap)
        return n.value, e[:k].copy(), w[:k].copy()

    def debug_phase_times(self, depth):
        p, mem, keep = _depth_arg(depth)
        out = (ctypes.c_int64 * 5)()
        self._ck(self.L.tsdf_debug_phase_times(self.h, p, mem, out))
        return [int(x) for x in out]

    def flush_l2(self):
        self._ck(self.L.tsdf_flush_l2(self.h))

    def ipc_export(self):
        buf = np.zeros(IPC_HANDLE_BYTES, np.uint8)
        self._ck(self.L.tsdf_shard_ipc_export(self.h, buf.ctypes.data_as(c_u8p)))
        return buf

    def ipc_attach(self, handles):
        a = np.ascontiguousarray(handles, np.uint8).reshape(-1, IPC_HANDLE_BYTES)
        self._ck(self.L.tsdf_shard_ipc_attach(self.h, len(a), a.ctypes.data_as(c_u8p)))


def pinned_empty(shape, dtype=np.float32):
    """numpy array over cudaMallocHost memory (kept alive by the returned array's base)."""
    L = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    st = L.tsdf_host_alloc_pinned(n, ctypes.byref(p))
    if st != 0:
        raise TsdfError(st, L.tsdf_last_error().decode())
    buf = (ctypes.c_uint8 * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


class ShardGroup:
    """z-slab shards living in this process (tsdf_shard_attach_local + tsdf_group_*)."""

    def __init__(self, n_shards, devices=None, bounds=None, **kw):
        self.L = load_library()
        devices = devices if devices is not None else [0] * n_shards
        ex = [dict(slab_k_begin=bounds[r], slab_k_end=bounds[r + 1]) if bounds is not None else {} for r in range(n_shards)]
        self.shards = [Tsdf(default_config(n_shards=n_shards, shard_rank=r, device=devices[r], **ex[r], **kw)) for r in range(n_shards)]
        self.n = n_shards
        self.arr = (ctypes.c_void_p * n_shards)(*[s.h for s in self.shards])
        self._ck(self.L.tsdf_shard_attach_local(self.arr, n_shards))
        self.m = self.shards[0].m

    def _ck(self, st):
        if st != 0:
            raise TsdfError(st, self.L.tsdf_last_error().decode())

    def close(self):
        for s in self.shards[::-1]:      # shard 0 owns the shared stream: destroy it last
            s.close()

    def set_intrinsics(self, K):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        self._ck(self.L.tsdf_group_set_intrinsics(self.arr, self.n, _d(K)))

    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float64).reshape(9); t = np.ascontiguousarray(t, np.float64).reshape(3)
        self._ck(self.L.tsdf_group_set_pose(self.arr, self.n, _d(R), _d(t)))

    def linearize(self, depth):
        p, mem, keep = _depth_arg(depth)
        A = np.empty(36); b = np.empty(6); st = TrackStats()
        self._ck(self.L.tsdf_group_linearize(self.arr, self.n, p, mem, _d(A), _d(b), ctypes.byref(st)))
        return A.reshape(6, 6), b, st.as_dict()

    def frame(self, depth, track=True, fuse=True):
        p, mem, keep = _depth_arg(depth)
        R = np.empty(9); t = np.empty(3); st = TrackStats(); n = ctypes.c_int64()
        self._ck(self.L.tsdf_group_frame(self.arr, self.n, p, mem, int(track), int(fuse), _d(R), _d(t), ctypes.byref(st), ctypes.byref(n)))
        return R.reshape(3, 3), t, st.as_dict(), n.value

    def download(self):
        """Assemble the full grid (reference layout [i,j,k]) from the slabs' OWNED layers."""
        m = self.m
        D = np.empty((m, m, m), np.float32); W = np.empty((m, m, m), np.float32)
        for s in self.shards:
            ks0, ks1, ko0, ko1 = s.stored_range()
            d, w = s.download(LAYOUT_REFERENCE)
            D[:, :, ko0:ko1] = d[:, :, ko0 - ks0:ko1 - ks0]
            W[:, :, ko0:ko1] = w[:, :, ko0 - ks0:ko1 - ks0]
        return D, W
