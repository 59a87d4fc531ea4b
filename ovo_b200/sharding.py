"""Multi-GPU layout of the hot path, one process per GPU over torch.distributed.

* Encoder / text tower: replicas.  Keyframe f of a batch is encoded by rank f % world; region descriptors are
  all-gathered (<= 0.2 MB per keyframe) so every map shard can fuse them.
* Map: points are partitioned by a spatial hash of their voxel (`shard_of_points`); a point never moves between
  shards.  Per keyframe every rank runs pass 1 on its own points (`vote`), the small vote tables
  [n_masks, n_instances+1] are summed with ONE all-reduce, and every rank takes the same decisions and updates its
  own points (`apply`).  The instance registry is therefore replicated deterministically without extra traffic.
* Query: row parallel, results stay sharded.

`ShardedAssociation` only needs an object with `vote(...) -> table` and `apply(table, ...)`: on the GPU that is
`ovo_b200.map.SemanticMap` (ovo_map_vote / ovo_map_apply, NCCL all-reduce); the CPU tests drive the same protocol
over gloo with a numpy backend."""
import numpy as np
import torch
import torch.distributed as dist


def shard_of_points(xyz, world: int, cell: float = 0.25):
    """Spatial hash of the voxel a point falls in -> shard id in [0, world).  Works on numpy arrays and torch
    tensors (any device); deterministic across ranks."""
    if torch.is_tensor(xyz):
        v = torch.floor(xyz / cell).to(torch.int64)
        h = (v[:, 0] * 73856093) ^ (v[:, 1] * 19349663) ^ (v[:, 2] * 83492791)
        return (h % world + world) % world
    v = np.floor(np.asarray(xyz, np.float64) / cell).astype(np.int64)
    h = (v[:, 0] * 73856093) ^ (v[:, 1] * 19349663) ^ (v[:, 2] * 83492791)
    return (h % world + world) % world


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Keyframes whose descriptors this rank computes (frame f -> rank f % world)."""
    return [f for f in range(n_frames) if f % world == rank]


class ShardedAssociation:
    """Association of one keyframe against a map sharded over the ranks of `group`."""

    def __init__(self, backend, group=None):
        self.backend, self.group = backend, group

    def associate(self, *vote_args, next_ins_id: int, **vote_kwargs):
        table = self.backend.vote(*vote_args, n_ins=next_ins_id, **vote_kwargs)       # local points only
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(table, op=dist.ReduceOp.SUM, group=self.group)             # the one exchange per keyframe
        return self.backend.apply(table, next_ins_id=next_ins_id)


def gather_descriptors(local_feats: torch.Tensor, counts_per_rank, group=None) -> torch.Tensor:
    """All-gather of the region descriptors computed by each rank (variable row counts) -> [sum, D] on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_feats
    D = local_feats.shape[1]
    mx = max(counts_per_rank)
    pad = torch.zeros(mx, D, dtype=local_feats.dtype, device=local_feats.device)
    pad[: local_feats.shape[0]] = local_feats
    out = [torch.empty_like(pad) for _ in counts_per_rank]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts_per_rank)])
