"""Colour fusion timings alone (dense 48 B/voxel pass and one trajectory frame); run on the GPU box."""
import sys, json
sys.path.insert(0, ".")
import bench, tracking_sdf_b200 as T
from tools import synth
depth, Rs, ts = synth.render_sequence(12)
print(json.dumps(bench.color_fuse_bench(T, 512, synth.K_DEFAULT, 5, bench.peaks()[0], depth[10], Rs[10], ts[10])))
