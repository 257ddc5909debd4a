"""Run the dense fusion micro-benchmark alone (for ncu captures)."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
import bench, tracking_sdf_b200 as T
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
print(json.dumps(bench.dense_fuse_bench(T, m, synth.K_DEFAULT, reps, bench.peaks()[0])))
