"""Host-side logic of z-slab sharding across processes (one process per GPU).

Every rank creates a handle with n_shards = world, shard_rank = rank (tsdf_create stores and fuses
its slab + halo), the ranks exchange the CUDA-IPC handles of their 30-double mailboxes
(torch.distributed is used only for that plumbing) and attach them; from then on the only
inter-GPU traffic is inside the tracking kernel: each rank's last block stores its partial
normal equations into every peer's mailbox over NVLink and sums all of them in rank order.
Fusion needs no exchange: halo layers are fused redundantly from the replicated depth frame.
"""
import numpy as np

from . import capi


def plan(m, world, **cfg_kw):
    """Slab plan of every rank (pure host computation)."""
    return [capi.slab_plan(capi.default_config(m=m, n_shards=world, shard_rank=r, **cfg_kw)) for r in range(world)]


def frustum_weights(m, K, poses, depths, extents=(6.0, 6.0, 3.5), origin=(-3.0, -3.0, -0.5), delta=0.3, coarse=64,
                    w_front=1.0, w_behind=0.25, w_floor=0.02):
    """Per-layer fusion cost profile for tsdf_balanced_slabs: on a coarse^3 sampling of the volume, for each
    representative (R, t, depth) count the voxels that project into the image with z > 0, weighted w_front when
    they are in front of the observed surface + delta (updated: exact or free-space path) and w_behind when
    behind it (classified and skipped), plus a floor for the per-row planning work.  Returns m weights."""
    K = np.asarray(K, float).reshape(3, 3)
    c = (np.arange(coarse) + 0.5) / coarse
    gx, gy, gz = np.meshgrid(c * extents[0] + origin[0], c * extents[1] + origin[1], c * extents[2] + origin[2], indexing="ij")
    pts = np.stack([gx, gy, gz], -1).reshape(-1, 3)
    w_layer = np.full(coarse, w_floor * coarse * coarse * len(poses), float)
    for (R, t), depth in zip(poses, depths):
        R = np.asarray(R, float); t = np.asarray(t, float)
        cam = (pts - t) @ R                      # R^T (p - t)
        z = cam[:, 2]
        ok = z > 0.05
        u = np.where(ok, K[0, 0] * cam[:, 0] / np.where(ok, z, 1.0) + K[0, 2], -1.0)
        v = np.where(ok, K[1, 1] * cam[:, 1] / np.where(ok, z, 1.0) + K[1, 2], -1.0)
        h, w = depth.shape
        ok &= (u >= 0) & (u < w) & (v >= 0) & (v < h)
        iu = np.clip(u.astype(np.int64), 0, w - 1); iv = np.clip(v.astype(np.int64), 0, h - 1)
        d = depth[iv, iu]
        ok &= np.isfinite(d)
        front = ok & (z <= d + delta)
        cost = (w_front * front + w_behind * (ok & ~front)).reshape(coarse, coarse, coarse)
        w_layer += cost.sum(axis=(0, 1))
    # coarse layers -> m layers (piecewise constant)
    return np.repeat(w_layer, (m + coarse - 1) // coarse)[:m] if m >= coarse else w_layer[:: coarse // m][:m]


def gather_bytes(dist, payload: np.ndarray) -> np.ndarray:
    """All-gather a fixed-size uint8 payload; returns [world, len] in rank order."""
    import torch
    world = dist.get_world_size()
    mine = torch.from_numpy(np.ascontiguousarray(payload, np.uint8).copy())
    backend = dist.get_backend()
    if backend == "nccl":
        mine = mine.cuda()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.stack([o.cpu().numpy() for o in out])


class ShardedTsdf:
    """This rank's slab of a volume sharded over dist's world."""

    def __init__(self, dist, device, bounds=None, **cfg_kw):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if bounds is not None:                 # explicit (work-balanced) partition, identical on every rank
            cfg_kw = dict(cfg_kw, slab_k_begin=int(bounds[self.rank]), slab_k_end=int(bounds[self.rank + 1]))
        self.t = capi.Tsdf(capi.default_config(n_shards=self.world, shard_rank=self.rank, device=device, **cfg_kw))
        handles = gather_bytes(dist, self.t.ipc_export())
        self.t.ipc_attach(handles)
        dist.barrier()

    def __getattr__(self, name):          # set_intrinsics, set_pose, track_and_fuse, enqueue_frame, ...
        return getattr(self.t, name)

    def close(self):
        self.dist.barrier()
        self.t.close()
