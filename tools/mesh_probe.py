"""Time tsdf_mesh_extract at 512^3 on a volume fused from 40 trajectory frames (run on the GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth, Rs, ts = synth.render_sequence(40)
g = T.Tsdf(T.default_config(m=m)); g.set_intrinsics(synth.K_DEFAULT)
for f in range(0, 40, 4):
    g.fuse(depth[f], Rs[f], ts[f])
best = 1e9
for _ in range(5):
    t0 = time.perf_counter(); n = g.mesh_extract(0.0); best = min(best, time.perf_counter() - t0)
print("m %d: %d triangles, %.3f ms per extraction (%.0f GB/s of store reads)" % (m, n // 3, best * 1e3, 2 * 8.0 * m ** 3 / best / 1e9))
