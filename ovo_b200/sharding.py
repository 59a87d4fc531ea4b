"""Multi-GPU layout of the hot path, one process per GPU over torch.distributed.

* Encoder / text tower: replicas.  Keyframe f of a batch is encoded by rank f % world; region descriptors are
  all-gathered (<= 0.2 MB per keyframe) so every map shard can fuse them.
* Map: points are partitioned by a spatial hash of their voxel (`shard_of_points`); a point never moves between
  shards.  Per keyframe every rank runs pass 1 on its own points (`vote`), the small vote tables
  [n_masks, n_instances+1] are summed with ONE all-reduce, and every rank takes the same decisions and updates its
  own points (`apply`).  The instance registry is therefore replicated deterministically without extra traffic.
* Map growth: the rank that integrates a frame (`ovo_map_integrate`, VanillaMapper.map) creates points whose voxels may
  belong to other shards; `route_new_points` sends each new point to its owner with ONE all-to-all per mapped frame
  (20 B per point: xyz f32x3, id i32, rgb u8x3 — <= 76.8k points ~ 1.5 MB), the only data-path exchange besides the vote
  table.  The receive order is deterministic (by source rank, then creation order).
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


class ShardedBatchAssociation:
    """A BATCH of keyframes against a map sharded over the ranks of `group` (ovo_map_batch_*): one pass over this rank's
    points for all keyframes, then per keyframe (in order) votes -> SUM of the vote tables over the ranks -> decisions; the
    instance count stays on the device, so the whole batch costs one host synchronisation.  `backend` needs
    batch_begin / batch_vote(f) -> table view / batch_decide(f) / batch_end (ovo_b200.map.SemanticMap; the CPU tests use a
    numpy backend).  `exchange`: "nccl" = one all-reduce per keyframe on the current stream; a callable(table, f) replaces it
    (the fused device-side exchange of ovo_b200.p2p)."""

    def __init__(self, backend, group=None, exchange="nccl"):
        self.backend, self.group, self.exchange = backend, group, exchange

    def associate(self, *begin_args, n_frames: int, mask_ins_out=None, **begin_kwargs):
        self.backend.batch_begin(*begin_args, **begin_kwargs)
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        for f in range(n_frames):
            table = self.backend.batch_vote(f)
            if multi:
                if callable(self.exchange):
                    self.exchange(table, f)
                else:
                    dist.all_reduce(table, op=dist.ReduceOp.SUM, group=self.group)     # the one exchange per keyframe
            self.backend.batch_decide(f)
        return self.backend.batch_end(mask_ins_out=mask_ins_out) if mask_ins_out is not None else self.backend.batch_end()


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


def route_new_points(xyz: torch.Tensor, ids: torch.Tensor, colors: torch.Tensor | None = None, group=None, cell: float = 0.25):
    """All-to-all of freshly created map points to the shards that own their voxels.
    xyz [n,3] f32, ids [n] i32, colors [n,3] u8 (optional), all on this rank's device (or CPU under gloo).
    Returns (xyz, ids, colors) of the points this rank now owns, ordered by source rank then by creation order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return xyz, ids, colors
    world = dist.get_world_size(group)
    dst = shard_of_points(xyz, world, cell)
    order = torch.argsort(dst, stable=True)
    send_counts = torch.bincount(dst, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    # one 20-byte record per point: 3 x f32 coordinates, the id and the packed colour reinterpreted as f32 lanes
    rec = torch.empty(xyz.shape[0], 5, dtype=torch.float32, device=xyz.device)
    rec[:, :3] = xyz
    rec[:, 3] = ids.to(torch.int32).view(torch.float32)
    if colors is not None:
        c = colors.to(torch.int32)
        rec[:, 4] = (c[:, 0] | (c[:, 1] << 8) | (c[:, 2] << 16)).view(torch.float32)
    else:
        rec[:, 4] = 0
    rec = rec[order].contiguous()
    n_in, n_out = send_counts.tolist(), recv_counts.tolist()
    out = torch.empty(sum(n_out), 5, dtype=torch.float32, device=xyz.device)
    dist.all_to_all_single(out, rec, output_split_sizes=n_out, input_split_sizes=n_in, group=group)
    oxyz = out[:, :3].contiguous()
    oids = out[:, 3].contiguous().view(torch.int32)
    ocol = None
    if colors is not None:
        p = out[:, 4].contiguous().view(torch.int32)
        ocol = torch.stack([p & 255, (p >> 8) & 255, (p >> 16) & 255], dim=1).to(torch.uint8)
    return oxyz, oids, ocol


def route_new_points_fixed(xyz: torch.Tensor, ids: torch.Tensor, cap_per_dst: int, group=None, cell: float = 0.25,
                           far: float = 1.0e6):
    """route_new_points without a host synchronisation: every rank sends exactly `cap_per_dst` 16-byte records (xyz f32x3,
    id i32) to every rank — its points of that shard in creation order, the rest padded with a far-away sentinel (id -1) that
    no frustum ever contains.  Returns (xyz [world*cap,3], ids [world*cap], overflow flag tensor: > 0 if some shard received
    more than cap_per_dst points from one source and dropped the surplus — size cap_per_dst with a margin over n/world)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if xyz.is_cuda and world <= 16:
        # one kernel (ovo_route_pack): hash, stable per-shard positions, records — instead of ~15 torch launches
        from . import _lib
        xyz_c, ids_c = xyz.to(torch.float32).contiguous(), ids.to(torch.int32).contiguous()
        rec = torch.empty(world * cap_per_dst, 4, dtype=torch.float32, device=xyz.device)
        overflow = torch.zeros((), dtype=torch.int32, device=xyz.device)
        _lib.check(_lib.lib().ovo_route_pack(_lib.ptr(xyz_c), _lib.ptr(ids_c), xyz_c.shape[0], world, float(cell), int(cap_per_dst), float(far),
                                             _lib.ptr(rec), _lib.ptr(overflow), _lib.stream_ptr(xyz.device)), "ovo_route_pack")
    else:
        dst = shard_of_points(xyz, world, cell)
        order = torch.argsort(dst, stable=True)
        sdst = dst[order]
        first = torch.searchsorted(sdst, torch.arange(world, device=xyz.device, dtype=sdst.dtype))
        pos = torch.arange(xyz.shape[0], device=xyz.device) - first[sdst]            # position inside its destination's run
        keep = pos < cap_per_dst
        overflow = (~keep).sum()
        rec = torch.full((world * cap_per_dst, 4), far, dtype=torch.float32, device=xyz.device)
        rec[:, 3] = torch.full((world * cap_per_dst,), -1, dtype=torch.int32, device=xyz.device).view(torch.float32)
        slot = (sdst * cap_per_dst + pos)[keep]
        src = order[keep]
        rec[slot, :3] = xyz[src]
        rec[slot, 3] = ids[src].to(torch.int32).view(torch.float32)
    if world > 1:
        out = torch.empty_like(rec)
        dist.all_to_all_single(out, rec, group=group)
    else:
        out = rec
    return out[:, :3].contiguous(), out[:, 3].contiguous().view(torch.int32), overflow
