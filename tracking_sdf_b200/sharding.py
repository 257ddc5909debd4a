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

    def __init__(self, dist, device, **cfg_kw):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.t = capi.Tsdf(capi.default_config(n_shards=self.world, shard_rank=self.rank, device=device, **cfg_kw))
        handles = gather_bytes(dist, self.t.ipc_export())
        self.t.ipc_attach(handles)
        dist.barrier()

    def __getattr__(self, name):          # set_intrinsics, set_pose, track_and_fuse, enqueue_frame, ...
        return getattr(self.t, name)

    def close(self):
        self.dist.barrier()
        self.t.close()
