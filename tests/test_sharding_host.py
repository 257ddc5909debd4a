"""Host-side logic of the multi-process sharded path, on CPU with the gloo backend (world size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from tracking_sdf_b200 import sharding, capi
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# 1. IPC-handle exchange plumbing: fixed-size byte payloads come back in rank order
payload = np.full(capi.IPC_HANDLE_BYTES, rank + 1, np.uint8); payload[0] = 100 + rank
got = sharding.gather_bytes(dist, payload)
assert got.shape == (world, capi.IPC_HANDLE_BYTES)
for r in range(world):
    assert got[r, 0] == 100 + r and (got[r, 1:] == r + 1).all()
# 2. every rank derives the same global slab plan and owns a distinct, contiguous part of it
pl = sharding.plan(1024, world)
mine = capi.slab_plan(capi.default_config(m=1024, n_shards=world, shard_rank=rank))
assert mine == pl[rank]
assert pl[0]["own"][0] == 0 and pl[-1]["own"][1] == 1024
assert all(pl[r]["own"][1] == pl[r + 1]["own"][0] for r in range(world - 1))
assert all(p["stored"][0] <= p["own"][0] and p["stored"][1] >= p["own"][1] for p in pl)
# 3. pixel ownership is a partition: each centre cell k belongs to exactly one rank
k = np.arange(1024)
owner = np.zeros(1024, int)
for r, p in enumerate(pl):
    owner[(k >= p["own"][0]) & (k < p["own"][1])] += 1
assert (owner == 1).all()
# 4. no GPU here: creating the shard must fail loudly, not fall back
try:
    capi.Tsdf(capi.default_config(m=64, n_shards=world, shard_rank=rank))
    ok = capi.load_library().tsdf_device_count() > 0
except capi.TsdfError as e:
    ok = e.status == 3
assert ok
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%d-ok\n" % rank); sys.stdout.flush()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_halo_covers_the_tracking_stencil():
    """SURVEY.md hard part 5: the halo must cover the centre cell, +-v_h voxels and the rotational
    perturbation w_h * |p| along z for every point of the volume."""
    sys.path.insert(0, ROOT)
    from tracking_sdf_b200 import sharding
    for m in (256, 1024, 2048):
        for world in (2, 8):
            for p in sharding.plan(m, world):
                vz = 3.5 / m
                reach = 1 + 1 + int(np.ceil(0.01 * np.sqrt(36 + 36 + 3.5 ** 2) / vz))
                assert p["halo"] >= reach
