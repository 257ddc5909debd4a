"""Phase timing of the sharded linearise kernel (torchrun, one rank per GPU)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tracking_sdf_b200 import sharding
from tools import synth
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group(backend="nccl")
depth, Rs, ts = synth.render_sequence(8)
g = sharding.ShardedTsdf(dist, local, m=m, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
g.set_intrinsics(synth.K_DEFAULT); g.set_pose(Rs[0], ts[0]); g.fuse(depth[0])
for f in range(1, 6):
    g.track_and_fuse(depth[f])
for rep in range(3):
    dist.barrier()
    t = g.debug_phase_times(depth[6])
    print("rank %d: main loop %.2f us | to final block %.2f | reduce %.2f | exchange+update %.2f | total %.2f" % ((dist.get_rank(),) + tuple(
        (b - a) / 1e3 for a, b in [(t[0], t[1]), (t[1], t[2]), (t[2], t[3]), (t[3], t[4]), (t[0], t[4])])), flush=True)
# ten chained iterations (one tsdf_track call) after a barrier: the track stage as the device saw it
for rep in range(3):
    g.set_pose(Rs[5], ts[5])
    dist.barrier(); torch.cuda.synchronize()
    g.track(depth[6])
    print("rank %d: track() of 10 chained iterations: %.1f us" % (dist.get_rank(), 1e3 * g.last_stage_ms()[1]), flush=True)
g.close(); dist.destroy_process_group()
