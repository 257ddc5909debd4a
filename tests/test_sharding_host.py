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
# 3b. the work-balanced partition: every rank derives identical cuts from the same cost profile, the explicit
#     slabs tile [0, m) and each rank's plan reports its own cut
wts = np.exp(-((np.arange(1024) - 400.0) / 150.0) ** 2) + 0.05
bounds = capi.balanced_slabs(wts, world, min_layers=16, halo=30)
allb = sharding.gather_bytes(dist, np.array(bounds, np.int32).view(np.uint8))
assert (allb == allb[0]).all()
assert bounds[0] == 0 and bounds[-1] == 1024 and all(bounds[r + 1] - bounds[r] >= 16 for r in range(world))
mine = capi.slab_plan(capi.default_config(m=1024, n_shards=world, shard_rank=rank, slab_k_begin=bounds[rank], slab_k_end=bounds[rank + 1]))
assert mine["own"] == (bounds[rank], bounds[rank + 1])
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
    """SURVEY.md hard part 5: the halo must cover the centre cell, +-v_h voxels and the rotational perturbation.
    (I +- w_h [e_k]x) rot moves a sample by w_h * e_k x v, v = R p = point - camera centre: its z component is at
    most w_h * max(|v_x|, |v_y|) <= w_h * max(width, height) while camera and point are over the volume's footprint.
    Checked numerically against random poses and points."""
    sys.path.insert(0, ROOT)
    from tracking_sdf_b200 import sharding
    rng = np.random.default_rng(11)
    w_h = 0.01
    worst = 0.0
    for _ in range(2000):
        cam = rng.uniform([-3, -3, -0.5], [3, 3, 3.0]); pt = rng.uniform([-3, -3, -0.5], [3, 3, 3.0])
        v = pt - cam
        for k in range(3):
            e = np.zeros(3); e[k] = 1.0
            worst = max(worst, abs(w_h * np.cross(e, v)[2]))
    assert worst <= w_h * 6.0
    for m in (256, 1024, 2048):
        for world in (2, 8):
            for p in sharding.plan(m, world):
                vz = 3.5 / m
                reach = 1 + 1 + int(np.ceil(worst / vz))
                assert p["halo"] >= reach


def test_balanced_slabs_minimise_the_largest_cost():
    """tsdf_balanced_slabs: cuts tile [0, m), respect min_layers, and the largest halo-inclusive slab cost is no
    worse than equal slabs' and within a layer's weight of the brute-force optimum on a small case."""
    import itertools
    from tracking_sdf_b200 import capi
    rng = np.random.default_rng(3)
    m, n, halo, minl = 40, 4, 2, 3
    w = rng.uniform(0.0, 1.0, m) ** 3 + 0.01
    P = np.concatenate([[0.0], np.cumsum(w)])

    def worst(b):
        return max(P[min(b[r + 1] + halo, m)] - P[max(b[r] - halo, 0)] for r in range(n))

    b = capi.balanced_slabs(w, n, min_layers=minl, halo=halo)
    assert b[0] == 0 and b[-1] == m and all(b[r + 1] - b[r] >= minl for r in range(n))
    best = min(worst((0,) + c + (m,)) for c in itertools.combinations(range(1, m), n - 1)
               if all(y - x >= minl for x, y in zip((0,) + c, c + (m,))))
    assert worst(b) <= best + 1e-9
    eq = [m * r // n for r in range(n + 1)]
    assert worst(b) <= worst(eq) + 1e-12
    # uniform weights, no halo -> equal slabs
    assert capi.balanced_slabs(np.ones(64), 4, min_layers=1, halo=0) == [0, 16, 32, 48, 64]
    # errors are reported, not guessed around
    import pytest
    with pytest.raises(capi.TsdfError):
        capi.balanced_slabs(np.ones(8), 4, min_layers=3)
    with pytest.raises(capi.TsdfError):
        capi.balanced_slabs(np.array([1.0, np.nan, 1.0, 1.0]), 2, min_layers=1)


def test_frustum_weights_profile():
    from tracking_sdf_b200 import sharding
    from tools import synth
    depth, Rs, ts = synth.render_sequence(2)
    w = sharding.frustum_weights(256, synth.K_DEFAULT, [(Rs[0], ts[0])], [depth[0]])
    assert w.shape == (256,) and (w > 0).all()
    # the camera sits at z ~ 1.3 m looking slightly down: the layers around it carry more than the top of the volume
    assert w[100:160].mean() > 3 * w[-20:].mean()


def test_pixel_aware_balancing():
    """tsdf_balanced_slabs2: a second profile that counts for a layer's OWNER only (the tracked pixels).  With no
    second profile it equals tsdf_balanced_slabs; with one, the largest (halo-inclusive fusion + owned pixels) cost is
    no larger than for the fusion-only cuts, and matches brute force on a small case."""
    import itertools
    import numpy as np
    from tracking_sdf_b200 import capi, sharding
    rng = np.random.default_rng(3)
    m, n, halo, minl = 48, 4, 3, 4
    wf = rng.random(m) + 0.05
    wo = np.zeros(m); wo[20:28] = 3.0                                 # the surface band sits in a few layers
    assert capi.balanced_slabs(wf, n, min_layers=minl, halo=halo) == capi.balanced_slabs(wf, n, min_layers=minl, halo=halo, weights_own=np.zeros(m))
    P = np.concatenate([[0], np.cumsum(wf)]); Q = np.concatenate([[0], np.cumsum(wo)])

    def worst(b):
        return max((P[min(b[r + 1] + halo, m)] - P[max(b[r] - halo, 0)]) + (Q[b[r + 1]] - Q[b[r]]) for r in range(n))
    b_f = capi.balanced_slabs(wf, n, min_layers=minl, halo=halo)
    b_p = capi.balanced_slabs(wf, n, min_layers=minl, halo=halo, weights_own=wo)
    assert b_p[0] == 0 and b_p[-1] == m and all(b_p[r + 1] - b_p[r] >= minl for r in range(n))
    best = min(worst((0,) + c + (m,)) for c in itertools.combinations(range(minl, m - minl + 1), n - 1)
               if all(y - x >= minl for x, y in zip((0,) + c, c + (m,))))
    assert worst(b_p) <= worst(b_f) + 1e-12 and abs(worst(b_p) - best) <= 1e-9 * best
    # the pixel profile itself: every tracked, valid pixel of a frame lands in exactly one layer
    from tools import synth
    depth, Rs, ts = synth.render_sequence(2)
    w = sharding.pixel_weights(64, synth.K_DEFAULT, [(Rs[0], ts[0])], [depth[0]])
    assert w.sum() == 214 * 160 and (w >= 0).all()
