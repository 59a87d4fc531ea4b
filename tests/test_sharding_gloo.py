"""The sharded-map protocol (ovo_b200/sharding.py) on CPU: 2 ranks over gloo, a numpy backend built on the oracle.
Sharded association over several keyframes must give every point the id the unsharded oracle gives it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fusion as OF
from ovo_b200 import synth
from ovo_b200.sharding import ShardedAssociation, shard_of_points, gather_descriptors, frames_of_rank


class NumpyBackend:
    """vote/apply with the oracle's arithmetic (stands in for SemanticMap on machines without a GPU)."""

    def vote(self, xyz, ins, depth, seg, c2w, K, n_ins, n_masks):
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        self.seg_of_pt, _ = OF.associate(xyz, ins, depth, seg, c2w, w2c, K, 0.05, True) if len(xyz) else (np.zeros(0, np.int32), None)
        self.ins, self.seg = ins, seg
        t = OF.vote_table(ins, self.seg_of_pt, n_masks, n_ins)
        return torch.from_numpy(np.concatenate([t.reshape(-1), [(self.seg_of_pt > -2).sum()]]).astype(np.int32))

    def apply(self, table, next_ins_id):
        t = table.numpy()
        n_masks = int(self.seg.max()) + 1
        votes = t[:-1].reshape(n_masks, -1)
        areas = np.array([(self.seg == m).sum() for m in range(n_masks)])
        rows, nxt = OF.decide_from_table(votes, areas, 100, next_ins_id)
        return OF.apply_decisions(self.ins, self.seg_of_pt, rows), rows, int(t[-1]), nxt


def _scene():
    K = synth.intrinsics(); d0 = synth.depth_map(frame_id=0)
    xyz, ids, ins = synth.point_map(60000, d0, K, synth.pose(0), seed=3, frac_visible=0.6)
    frames = []
    for i in range(3):
        seg, _ = synth.grid_masks(rows=(6 if i % 2 == 0 else 3), cols=(8 if i % 2 == 0 else 5))
        frames.append((synth.depth_map(frame_id=4 * i), seg, synth.pose(4 * i)))
    return K, xyz, ins, frames


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, xyz, ins, frames = _scene()
    mine = shard_of_points(xyz, world) == rank
    lxyz, lins = xyz[mine], ins[mine]
    sa = ShardedAssociation(NumpyBackend())
    nxt, log = 0, []
    for depth, seg, c2w in frames:
        lins, rows, n_matched, nxt = sa.associate(lxyz, lins, depth, seg, c2w, K, n_masks=int(seg.max()) + 1, next_ins_id=nxt)
        log.append(([r["ins_id"] for r in rows], n_matched, nxt))
    feats = torch.full((2 + rank, 4), float(rank))
    allf = gather_descriptors(feats, [2 + r for r in range(world)])
    q.put((rank, np.nonzero(mine)[0], lins, log, allf.numpy(), frames_of_rank(5, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_sharded_association_matches_unsharded_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unsharded oracle
    K, xyz, ins, frames = _scene()
    nxt, ref_log = 0, []
    for depth, seg, c2w in frames:
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        sp, _ = OF.associate(xyz, ins, depth, seg, c2w, w2c, K, 0.05, True)
        ins, rows, nxt = OF.track(ins, sp, seg, 100, nxt)
        ref_log.append(([r["ins_id"] for r in rows], int((sp > -2).sum()), nxt))
    merged = np.full(len(xyz), -99, np.int32)
    for rank, idx, lins, log, allf, fr in res:
        merged[idx] = lins
        assert log == ref_log                                      # every rank took the same decisions as the unsharded run
        assert allf.shape == (5, 4) and allf[:2].max() == 0 and allf[2:].min() == 1
        assert fr == [f for f in range(5) if f % world == rank]
    assert (merged == ins).all()
    assert nxt > 40


def test_shard_function_is_deterministic_and_balanced():
    xyz = np.random.default_rng(0).uniform(-8, 8, (200000, 3)).astype(np.float32)
    a = shard_of_points(xyz, 8)
    b = shard_of_points(torch.from_numpy(xyz), 8).numpy()
    assert (a == b).all() and a.min() == 0 and a.max() == 7
    counts = np.bincount(a, minlength=8)
    assert counts.min() > 0.8 * counts.mean()


def test_vote_table_decomposition_equals_track():
    K, xyz, ins, frames = _scene()
    depth, seg, c2w = frames[0]
    w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
    sp, _ = OF.associate(xyz, ins, depth, seg, c2w, w2c, K, 0.05, True)
    ins1, rows1, n1 = OF.track(ins, sp, seg, 100, 0)
    n_masks = int(seg.max()) + 1
    rows2, n2 = OF.decide_from_table(OF.vote_table(ins, sp, n_masks, 0), [r["area"] for r in rows1], 100, 0)
    assert rows1 == rows2 and n1 == n2 and (OF.apply_decisions(ins, sp, rows2) == ins1).all()
    # second keyframe, with assigned points present
    depth, seg, c2w = frames[1]
    w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
    sp, _ = OF.associate(xyz, ins1, depth, seg, c2w, w2c, K, 0.05, True)
    ins2, rows1, m1 = OF.track(ins1, sp, seg, 100, n1)
    n_masks = int(seg.max()) + 1
    rows2, m2 = OF.decide_from_table(OF.vote_table(ins1, sp, n_masks, n1), [r["area"] for r in rows1], 100, n1)
    assert rows1 == rows2 and m1 == m2 and (OF.apply_decisions(ins1, sp, rows2) == ins2).all()


def _route_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ovo_b200.sharding import route_new_points
    rng = np.random.default_rng(100 + rank)                       # each rank integrated a different frame
    n = 5000 + 700 * rank
    xyz = torch.from_numpy(rng.uniform(-4, 4, (n, 3)).astype(np.float32))
    ids = torch.arange(n, dtype=torch.int32) + 1_000_000 * rank
    col = torch.from_numpy(rng.integers(0, 256, (n, 3), dtype=np.uint8))
    oxyz, oids, ocol = route_new_points(xyz, ids, col)
    q.put((rank, xyz.numpy(), ids.numpy(), col.numpy(), oxyz.numpy(), oids.numpy(), ocol.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_new_points_are_routed_to_their_shard_world2():
    """The one all-to-all of the map-growth step: every new point ends on the shard its voxel hashes to,
    nothing is lost or duplicated, payload (xyz bits, id, colour) intact, receive order = (source rank, creation order)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_route_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *_rest in res:
        oxyz, oids, ocol = res[rank][4], res[rank][5], res[rank][6]
        assert (shard_of_points(oxyz, world) == rank).all()
        exp_xyz, exp_ids, exp_col = [], [], []
        for src in range(world):                                   # source-rank major, creation order inside
            sx, si, sc = res[src][1], res[src][2], res[src][3]
            m = shard_of_points(sx, world) == rank
            exp_xyz.append(sx[m]); exp_ids.append(si[m]); exp_col.append(sc[m])
        assert (oxyz.view(np.int32) == np.concatenate(exp_xyz).view(np.int32)).all()      # bit-exact coordinates
        assert (oids == np.concatenate(exp_ids)).all() and (ocol == np.concatenate(exp_col)).all()
    assert sum(len(r[5]) for r in res) == sum(len(r[2]) for r in res)


# ---------------------------------------------------------------------------------------------- batch protocol (round 2)
class NumpyBatchBackend:
    """batch_begin / batch_vote / batch_decide / batch_end (ovo_map_batch_*) with the oracle's arithmetic: the geometry of every
    keyframe first, then per keyframe votes on the current ids -> (summed) table -> decisions applied to this shard."""

    def batch_begin(self, xyz, ins, frames, K, next_ins_id):
        self.ins, self.frames, self.nxt = ins.copy(), frames, next_ins_id
        self.segs = []
        for depth, seg, c2w in frames:
            w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
            sp, _ = OF.associate(xyz, ins, depth, seg, c2w, w2c, K, 0.05, True) if len(xyz) else (np.zeros(0, np.int32), None)
            self.segs.append(sp)
        self.rows = []

    def batch_vote(self, f):
        seg = self.frames[f][1]
        self.n_masks = int(seg.max()) + 1
        t = OF.vote_table(self.ins, self.segs[f], self.n_masks, self.nxt)
        self.table = torch.from_numpy(np.concatenate([t.reshape(-1), [(self.segs[f] > -2).sum()]]).astype(np.int32))
        return self.table

    def batch_decide(self, f):
        seg = self.frames[f][1]
        t = self.table.numpy()
        areas = np.array([(seg == m).sum() for m in range(self.n_masks)])
        rows, self.nxt = OF.decide_from_table(t[:-1].reshape(self.n_masks, -1), areas, 100, self.nxt)
        self.ins = OF.apply_decisions(self.ins, self.segs[f], rows)
        self.rows.append((rows, int(t[-1])))

    def batch_end(self):
        return self.rows, self.ins, self.nxt


def _batch_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ovo_b200.sharding import ShardedBatchAssociation, route_new_points_fixed
    K, xyz, ins, frames = _scene()
    mine = shard_of_points(xyz, world) == rank
    sa = ShardedBatchAssociation(NumpyBatchBackend())
    rows, lins, nxt = sa.associate(xyz[mine], ins[mine], frames, K, 0, n_frames=len(frames))
    # fixed-size routing of new points (no host synchronisation: padded records)
    rng = np.random.default_rng(200 + rank)
    n = 3000
    nx = torch.from_numpy(rng.uniform(-4, 4, (n, 3)).astype(np.float32))
    ni = torch.arange(n, dtype=torch.int32) + 10_000 * rank
    ox, oi, ovf = route_new_points_fixed(nx, ni, cap_per_dst=2000)
    ox2, oi2, ovf2 = route_new_points_fixed(nx, ni, cap_per_dst=1000)     # too small on purpose: the surplus is counted
    q.put((rank, np.nonzero(mine)[0], lins, [([r["ins_id"] for r in rr], nm) for rr, nm in rows], nxt, nx.numpy(), ni.numpy(),
           ox.numpy(), oi.numpy(), int(ovf), int(ovf2), int((oi2 >= 0).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_association_and_fixed_routing_world2():
    """The batched protocol (one geometry pass, per keyframe vote -> all-reduce -> decide) gives every shard the ids and rows
    of the unsharded oracle; route_new_points_fixed delivers every new point to its owner, padded with sentinels."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    K, xyz, ins, frames = _scene()
    nxt, ref = 0, []
    for depth, seg, c2w in frames:
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        sp, _ = OF.associate(xyz, ins, depth, seg, c2w, w2c, K, 0.05, True)
        ins, rows, nxt = OF.track(ins, sp, seg, 100, nxt)
        ref.append(([r["ins_id"] for r in rows], int((sp > -2).sum())))
    merged = np.full(len(xyz), -99, np.int32)
    for rank, idx, lins, log, n, *_ in res:
        merged[idx] = lins
        assert log == ref and n == nxt
    assert (merged == ins).all()
    for rank, *_ in res:
        ox, oi, ovf, ovf2, kept2 = res[rank][7], res[rank][8], res[rank][9], res[rank][10], res[rank][11]
        assert ovf == 0 and ovf2 > 0 and kept2 <= world * 1000
        real = oi >= 0
        assert (shard_of_points(ox[real], world) == rank).all() and (ox[~real] > 1e5).all()
        exp = np.concatenate([res[s][6][shard_of_points(res[s][5], world) == rank] for s in range(world)])
        assert (oi[real] == exp).all()                           # source-rank major, creation order inside
    assert sum(int((r[8] >= 0).sum()) for r in res) == sum(len(r[6]) for r in res)
