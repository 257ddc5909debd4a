"""Multi-GPU correctness check of the z-slab sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_check.py [m] [frames]

Every rank owns one slab (+halo), tracking exchanges the normal equations in-kernel over NVLink
peer stores.  Each rank also runs the UNSHARDED volume on its own GPU and compares: poses per frame
(<= 1e-9), and its whole stored slab incl. the redundantly fused halo, bit for bit, after re-fusing
from identical poses.  Prints one line per rank; exit code 0 only if all comparisons hold."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import tracking_sdf_b200 as T
from tracking_sdf_b200 import sharding
from tools import synth


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    depth, Rs, ts = synth.render_sequence(nf)
    kw = dict(m=m, gauss_newton_max_iteration=10, maximum_twist_diff=float("-inf"))
    sh = sharding.ShardedTsdf(dist, local, **kw)
    one = T.Tsdf(T.default_config(device=local, **kw))
    for x in (sh, one):
        x.set_intrinsics(synth.K_DEFAULT); x.set_pose(Rs[0], ts[0])
    n_sh = sh.fuse(depth[0]); n_one = one.fuse(depth[0])
    worst_t = worst_r = 0.0
    ok = True
    for f in range(1, nf):
        R1, t1, s1, _ = one.track_and_fuse(depth[f])
        R2, t2, s2, _ = sh.track_and_fuse(depth[f])
        worst_t = max(worst_t, float(np.abs(t1 - t2).max())); worst_r = max(worst_r, float(np.abs(R1 - R2).max()))
        ok &= s2["iterations"] == 10 and s2["n_valid"] == s1["n_valid"] and s2["halo_miss"] == 0
        one.set_pose(R2, t2)          # keep both on the same pose bits
    ok &= worst_t < 1e-9 and worst_r < 1e-9
    # all ranks hold the same pose bits (each solved the same summed system)
    R, t = sh.get_pose()
    buf = torch.tensor(np.concatenate([R.ravel(), t]), device="cuda")
    ref = buf.clone(); dist.broadcast(ref, 0)
    ok &= bool(torch.equal(buf, ref))
    # fusion from identical poses: slab (with halo) == the same layers of the unsharded grid
    one.reset(); sh.reset()
    for x in (one, sh):
        x.set_intrinsics(synth.K_DEFAULT)
    for f in range(min(nf, 3)):
        one.fuse(depth[f], Rs[f], ts[f]); sh.fuse(depth[f], Rs[f], ts[f])
    ks0, ks1, ko0, ko1 = sh.stored_range()
    D1, W1 = one.download(); D2, W2 = sh.download()
    same = np.array_equal(D1[:, :, ks0:ks1], D2) and np.array_equal(W1[:, :, ks0:ks1], W2)
    ok &= same
    print("rank %d/%d m=%d slab own [%d,%d) stored [%d,%d): pose diff %.2e / %.2e, slab bit-equal %s -> %s"
          % (rank, world, m, ko0, ko1, ks0, ks1, worst_t, worst_r, same, "OK" if ok else "FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sh.close(); one.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
