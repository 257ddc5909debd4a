"""K0 (pre-processing) on a few noisy frames at 512^3, for ncu launch lists; run on the GPU box."""
import sys
sys.path.insert(0, ".")
import tracking_sdf_b200 as T
from tools import synth
depth, Rs, ts = synth.render_sequence(4)
noisy = synth.add_sensor_noise(depth, seed=1234)
g = T.Tsdf(T.default_config(m=512, preprocess=1, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf")))
g.set_intrinsics(synth.K_DEFAULT)
g.fuse(noisy[0], Rs[0], ts[0])
for f in range(1, 4):
    g.track_and_fuse(noisy[f])
print("ok")
g.close()
