"""Certificate statistics of trajectory frames at 512^3 on the CPU emulation (no GPU needed)."""
import sys, ctypes
import numpy as np
sys.path.insert(0, ".")
from tests import emul
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
margin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
depth, Rs, ts = synth.render_sequence(40)
e = emul.Emul(synth.K_DEFAULT, m=m, metric=1)
L = e.L
names = ["items", "row FRONT", "item FRONT", "item SKIP", "per-lane items", " all-UNKNOWN", " no-UNKNOWN", "units UNKNOWN", "units FRONT", "units SKIP",
         "pred hopeless", " really hopeless", " certified units lost"]
for f in (0, 20, 39):
    e.set_pose(Rs[f], ts[f]); e.prep(depth[f])
    out = (ctypes.c_int64 * 16)()
    L.emul_fuse_stats(e.g, e.pix.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), e.pose, ctypes.c_float(margin), out)
    print("frame", f, " ".join("%s=%d" % (n, out[i]) for i, n in enumerate(names)))
