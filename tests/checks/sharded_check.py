"""Sharded-map association on N GPUs (NCCL): every rank owns the points whose voxel hashes to it, votes are summed
with one all-reduce per keyframe, decisions are identical on every rank.  Checked against the unsharded oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/checks/sharded_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fusion as OF  # noqa: E402
from ovo_b200 import synth  # noqa: E402
from ovo_b200.map import SemanticMap  # noqa: E402
from ovo_b200.sharding import ShardedAssociation, shard_of_points, gather_descriptors, route_new_points  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K = synth.intrinsics(); d0 = synth.depth_map(frame_id=0)
    N = 400000
    xyz, ids, ins = synth.point_map(N, d0, K, synth.pose(0), seed=3, frac_visible=0.6)
    mine = shard_of_points(xyz, world) == rank
    lxyz = torch.from_numpy(xyz[mine]).to(dev)
    lins = torch.from_numpy(ins[mine]).to(dev)
    sm = SemanticMap(dev)

    class Backend:
        def vote(self, depth, seg, c2w, n_ins, n_masks, slot):
            return sm.vote(lxyz, lins, depth, seg, c2w, K, n_ins=n_ins, n_masks=n_masks, kf_slot=slot)

        def apply(self, table, next_ins_id):
            return sm.apply(table, next_ins_id)

    sa = ShardedAssociation(Backend())
    nxt, ref_ins, ref_nxt, ok = 0, ins.copy(), 0, True
    for i in range(3):
        seg, bm = synth.grid_masks(rows=(6 if i % 2 == 0 else 3), cols=(8 if i % 2 == 0 else 5))
        depth, c2w = synth.depth_map(frame_id=4 * i), synth.pose(4 * i)
        votes, n_matched, nxt = sa.associate(torch.from_numpy(depth).to(dev), torch.from_numpy(seg).to(dev), c2w,
                                             n_masks=int(seg.max()) + 1, slot=i, next_ins_id=nxt)
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        sp, _ = OF.associate(xyz, ref_ins, depth, seg, c2w, w2c, K, 0.05, True)
        ref_ins, rows, ref_nxt = OF.track(ref_ins, sp, seg, 100, ref_nxt)
        same = (lins.cpu().numpy() == ref_ins[mine]).all() and nxt == ref_nxt and n_matched == int((sp > -2).sum()) and \
            (votes["ins_id"] == np.array([r["ins_id"] for r in rows])).all() and (votes["n_matched"] == np.array([r["n_matched"] for r in rows])).all()
        ok = ok and bool(same)
        if rank == 0:
            print(f"keyframe {i}: n_matched {n_matched} next_id {nxt} match_oracle {bool(same)}", flush=True)
    feats = torch.full((2 + rank, 8), float(rank), device=dev)
    allf = gather_descriptors(feats, [2 + r for r in range(world)])
    ok = ok and allf.shape[0] == sum(2 + r for r in range(world))
    # map growth: every rank "integrated" a different frame; the new points go to their owners with one NCCL all-to-all
    rng = np.random.default_rng(100 + rank)
    n_new = 76800
    nxyz = torch.from_numpy(rng.uniform(-4, 4, (n_new, 3)).astype(np.float32)).to(dev)
    nids = (torch.arange(n_new, dtype=torch.int32) + 1_000_000 * rank).to(dev)
    ncol = torch.from_numpy(rng.integers(0, 256, (n_new, 3), dtype=np.uint8)).to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    oxyz, oids, ocol = route_new_points(nxyz, nids, ncol)
    e1.record()
    torch.cuda.synchronize()
    owned = bool((shard_of_points(oxyz, world) == rank).all())
    tot = torch.tensor([oxyz.shape[0]], device=dev)
    dist.all_reduce(tot)
    ok = ok and owned and int(tot.item()) == n_new * world and bool((oids // 1_000_000 < world).all())
    if rank == 0:
        print(f"route_new_points: {n_new} new points per rank, all-to-all {e0.elapsed_time(e1):.3f} ms, owned {owned}, total {int(tot.item())}", flush=True)
    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_CHECK", "PASS" if t.item() == 1 else "FAIL", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
