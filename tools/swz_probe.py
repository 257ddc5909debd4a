"""Layout experiment (variant build libtsdf_swz.so): k_linearize on the x-fastest store vs a swizzled copy."""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth, Rs, ts = synth.render_sequence(12)
g = T.Tsdf(T.default_config(m=m, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf")))
g.set_intrinsics(synth.K_DEFAULT); g.set_pose(Rs[0], ts[0]); g.fuse(depth[0])
for f in range(1, 10):
    g.track_and_fuse(depth[f])
L = g.L
L.tsdf_debug_swizzle_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)]
out = (ctypes.c_float * 2)(); eq = ctypes.c_int32()
d = np.ascontiguousarray(depth[10])
for rep in range(3):
    st = L.tsdf_debug_swizzle_probe(g.h, d.ctypes.data, 0, 50, out, ctypes.byref(eq))
    print("status %d: x-fastest %.2f us/launch, swizzled %.2f us/launch, sums equal %d" % (st, out[0] * 1e3, out[1] * 1e3, eq.value))
g.close()
