"""ctypes binding of tests/host_emul/core_emul.cpp (HOST compile of the kernels' core; CPU tests only)."""
import ctypes

import numpy as np

from tools import build

c_dp = ctypes.POINTER(ctypes.c_double)
c_fp = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build.build_emul())
        _lib.emul_fuse.restype = ctypes.c_int64
        _lib.emul_last_fast_count.restype = ctypes.c_int64
        _lib.emul_weight_exp.restype = ctypes.c_double
        _lib.emul_weight_exp.argtypes = [ctypes.c_double]
        _lib.emul_trunc_f2i.argtypes = [ctypes.c_float]
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f(a):
    return a.ctypes.data_as(c_fp)


class Emul:
    """Holds a GridParams blob, a PoseState blob and an x-fastest interleaved {D,W} grid."""

    def __init__(self, K, m=64, width=6.0, height=6.0, depth=3.5, origin=(-3.0, -3.0, -0.5), delta=0.3, eps=0.025,
                 v_h=1.0, w_h=0.01, stride=3, metric=0, img_w=640, img_h=480, max_twist_diff=0.001, max_iter=20):
        L = lib()
        self.L = L
        self.m, self.img_w, self.img_h, self.stride = m, img_w, img_h, stride
        self.g = ctypes.create_string_buffer(L.emul_sizeof_params())
        self.pose = ctypes.create_string_buffer(L.emul_sizeof_pose())
        org = np.asarray(origin, np.float64)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        L.emul_params(m, ctypes.c_float(width), ctypes.c_float(height), ctypes.c_float(depth), _d(org),
                      ctypes.c_float(delta), ctypes.c_float(eps), ctypes.c_float(v_h), ctypes.c_float(w_h),
                      stride, metric, img_w, img_h, _d(K), ctypes.c_float(max_twist_diff), max_iter, self.g)
        self.grid = np.empty((m, m, m, 2), np.float32)      # [k, j, i, {D,W}]
        self.grid[..., 0] = np.float32(width) + np.float32(height) + np.float32(depth)
        self.grid[..., 1] = 0
        self.pix = np.empty((img_h, img_w, 4), np.float32)
        self.P = ((img_w + stride - 1) // stride) * ((img_h + stride - 1) // stride)

    def set_pose(self, R, t):
        R = np.ascontiguousarray(R, np.float64).reshape(9); t = np.ascontiguousarray(t, np.float64).reshape(3)
        self.L.emul_pose_set(self.pose, _d(R), _d(t))

    def get_pose(self):
        R = np.empty(9); t = np.empty(3); Ri = np.empty(9); ti = np.empty(3)
        self.L.emul_pose_get(self.pose, _d(R), _d(t), _d(Ri), _d(ti))
        return R.reshape(3, 3), t, Ri.reshape(3, 3), ti

    def prep(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        self.L.emul_prep(self.g, _f(depth), _f(self.pix))

    def cloud(self):
        c = np.empty((self.img_h, self.img_w, 3), np.float32); n = np.empty_like(c)
        self.L.emul_cloud(self.g, _f(self.pix), _f(c), _f(n))
        return c, n

    def fuse(self, use_clip=1, use_fast=1):
        n = self.L.emul_fuse(self.g, _f(self.grid), _f(self.pix), self.pose, use_clip, use_fast)
        self.last_fast = self.L.emul_last_fast_count()
        return n

    # colour (k_fuse_cert in queue_front mode + k_fuse_exact<colour>, k_sample_color) and the mesher (k_mc_sweep)
    def fuse_rgb(self, rgb, use_clip=1, use_fast=1):
        if not hasattr(self, "color"):
            self.color = np.empty((self.m, self.m, self.m, 4), np.float32)       # [k, j, i, {Color_W,R,G,B}]
            self.color[..., 0] = 0.0; self.color[..., 1:] = np.float32(0.4)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.L.emul_fuse_rgb.restype = ctypes.c_int64
        n = self.L.emul_fuse_rgb(self.g, _f(self.grid), _f(self.pix), self.pose, use_clip, use_fast, _f(self.color),
                                 rgb.ctypes.data_as(c_u8p))
        self.last_fast = self.L.emul_last_fast_count()
        return n

    def color_ref(self):
        """(Color_W, R, G, B) in the reference's [i, j, k] indexing."""
        return tuple(self.color[..., q].transpose(2, 1, 0) for q in range(4))

    def interpolate_color(self, gpts):
        gpts = np.ascontiguousarray(gpts, np.float64).reshape(-1, 3)
        out = np.empty((len(gpts), 4), np.float32)
        self.L.emul_interpolate_color(self.g, _f(self.color), ctypes.c_int64(len(gpts)), _d(gpts), _f(out))
        return out

    def mesh(self, iso=0.0, extents=(6.0, 6.0, 3.5)):
        self.L.emul_mesh.restype = ctypes.c_int64
        args = (self.g, ctypes.c_float(extents[0]), ctypes.c_float(extents[1]), ctypes.c_float(extents[2]), ctypes.c_float(iso), _f(self.grid))
        n = self.L.emul_mesh(*args, None)
        xyz = np.empty((n, 3), np.float32)
        if n:
            self.L.emul_mesh(*args, _f(xyz))
        return xyz

    def interpolate(self, pts):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        out = np.empty(len(pts), np.float32); ok = np.empty(len(pts), np.uint8)
        self.L.emul_interpolate(self.g, _f(self.grid), ctypes.c_int64(len(pts)), _d(pts), _f(out), ok.ctypes.data_as(c_u8p))
        return out, ok.astype(bool)

    def linearize(self):
        J = np.empty((self.P, 6), np.float32); psi = np.empty(self.P, np.float32); flag = np.empty(self.P, np.uint8)
        sums = np.empty(30)
        self.L.emul_linearize(self.g, _f(self.grid), _f(self.pix), self.pose, _f(J), _f(psi), flag.ctypes.data_as(c_u8p), _d(sums))
        return J, psi, flag, sums

    def gn_update(self, sums):
        sums = np.ascontiguousarray(sums, np.float64)
        self.L.emul_gn_update(self.g, self.pose, _d(sums))
        st = np.empty(4, np.int32); tw = np.empty(6)
        self.L.emul_pose_stats(self.pose, st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _d(tw))
        return {"iterations": int(st[0]), "stopped": int(st[1]), "singular": int(st[2])}, tw

    # grid views in the reference's [i,j,k] indexing
    @property
    def D(self):
        return self.grid[..., 0].transpose(2, 1, 0)

    @property
    def W(self):
        return self.grid[..., 1].transpose(2, 1, 0)

    def load_reference_layout(self, D, W):
        self.grid[..., 0] = np.asarray(D).transpose(2, 1, 0)
        self.grid[..., 1] = np.asarray(W).transpose(2, 1, 0)


def sums_to_Ab(sums):
    A = np.zeros((6, 6)); q = 0
    for r in range(6):
        for c in range(r, 6):
            A[r, c] = A[c, r] = sums[q]; q += 1
    return A, sums[21:27].copy()
