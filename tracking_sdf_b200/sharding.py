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


def pixel_weights(m, K, poses, depths, extent_z=3.5, origin_z=-0.5, stride=3):
    """Per-layer count of TRACKED pixels (every `stride`-th pixel, camera_tracking.cpp:162-163) whose back-projected
    point falls into the layer, summed over representative (R, t, depth): the tracker's ownership profile (a rank
    linearises the pixels whose centre cell it owns).  Returns m counts."""
    K = np.asarray(K, float).reshape(3, 3)
    w = np.zeros(m)
    for (R, t), depth in zip(poses, depths):
        R = np.asarray(R, float); t = np.asarray(t, float)
        d = np.asarray(depth, np.float64)[::stride, ::stride]
        v, u = np.meshgrid(np.arange(0, depth.shape[0], stride), np.arange(0, depth.shape[1], stride), indexing="ij")
        ok = np.isfinite(d) & (d > 0)
        x = (u - K[0, 2]) * d / K[0, 0]; y = (v - K[1, 2]) * d / K[1, 1]
        wz = R[2, 0] * x + R[2, 1] * y + R[2, 2] * d + t[2]
        k = np.floor((wz - origin_z) * m / extent_z - 0.5).astype(np.int64)
        ok &= (k >= 0) & (k < m)
        w += np.bincount(k[ok], minlength=m)[:m]
    return w


def frame_cost_weights(m, K, poses, depths, fuse_ms=1.0, track_px_ms=0.16, **kw):
    """(weights, weights_own) for capi.balanced_slabs in one unit (milliseconds of a frame): the fusion profile scaled
    to the single-GPU fusion time of the whole volume, the pixel profile scaled to the pixel-loop time of a whole
    frame's tracking (10 iterations x 16 us on one B200: the part of tracking that shrinks with ownership)."""
    wf = frustum_weights(m, K, poses, depths, **kw)
    wp = pixel_weights(m, K, poses, depths)
    wf = wf * (fuse_ms / max(wf.sum(), 1e-300))
    wp = wp * (track_px_ms / max(wp.sum(), 1e-300))
    return wf, wp


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
