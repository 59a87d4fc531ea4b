"""The sharded map through the public `OVO` API on TWO GPUs over NCCL (skipped below 2 GPUs): every rank holds the
points whose voxel hashes to it, makes the same calls with the same keyframes; instance ids, vote decisions, the replicated
instance bank and the query result equal the single-GPU run, and the dense bank shards equal the rows of the single-GPU bank.
Also the batched association with the fused device-side vote exchange (ovo_map_associate_batch_sharded) against the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import gen_golden as GG

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class _Logger:
    def log_ovo_stats(self, *a, **k):
        pass


class _TokTokenizer:
    def __call__(self, phrase):
        return torch.from_numpy(GG.QUERIES_TOK[int(phrase)][None])


def _run_ovo(rank, world, port, tmp, q):
    import torch.distributed as dist
    from ovo_b200 import OVO, CLIPGenerator
    from ovo_b200.encoder import random_state_dict
    from ovo_b200.sharding import shard_of_points
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    K, xyz, ids, ins, frames = GG.ovo_inputs()
    mdir = os.path.join(tmp, f"masks{world}_{rank}", "scene")
    os.makedirs(mdir, exist_ok=True)
    for f in frames:
        np.save(os.path.join(mdir, f"{f['frame_id']:04d}_seg_map_default.npy"), f["seg"])
        np.save(os.path.join(mdir, f"{f['frame_id']:04d}_bmap_default.npy"), f["bm"])
    cfg = GG.tiny_cfg()
    config = GG.ovo_config(os.path.dirname(mdir))
    config["dense_map"] = True
    config["shard_map"] = world > 1
    clip = CLIPGenerator(config["clip"], state_dict=random_state_dict(cfg, seed=0), tokenizer=_TokTokenizer(), encoder_config=cfg,
                         device=f"cuda:{rank}")
    ovo = OVO(config, _Logger(), scene_name="scene", cam_intrinsics=torch.from_numpy(K), clip_generator=clip, device=f"cuda:{rank}")
    mine = np.nonzero(shard_of_points(xyz, world) == rank)[0] if world > 1 else np.arange(len(xyz))
    dev = torch.device("cuda", rank)
    pts, pids, pins = (torch.from_numpy(a[mine]).to(dev) for a in (xyz, ids, ins))
    for f in frames:
        upd = ovo.detect_and_track_objects((f["frame_id"], f["image"], f["depth"], ()), (pts, pids, pins), torch.from_numpy(f["c2w"]))
        pins = upd
        ovo.compute_semantic_info()
    ovo.complete_semantic_info()
    sim = ovo.query(["0", "1", "2"])
    torch.cuda.synchronize()
    q.put(dict(rank=rank, mine=mine, ins=pins.cpu().numpy(), objects=list(ovo.objects.keys()), sim=sim.cpu().numpy(),
               clips=ovo.get_objs_clips().cpu().numpy(), dense=ovo._dense_bank[: len(mine)].float().cpu().numpy(),
               dense_lo=ovo._dense_bank_lo[: len(mine)].float().cpu().numpy(), counts=ovo._dense_counts[: len(mine)].cpu().numpy(),
               next_id=ovo.next_ins_id))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _spawn(fn, world, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=fn, args=(r, world, port, *args, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ovo_shard_map_on_two_gpus_equals_the_single_gpu_run(tmp_path):
    one = _spawn(_run_ovo, 1, str(tmp_path))[0]
    two = _spawn(_run_ovo, 2, str(tmp_path))
    merged = np.full(len(one["ins"]), -99, np.int32)
    for r in two:
        merged[r["mine"]] = r["ins"]
        assert r["objects"] == one["objects"] and r["next_id"] == one["next_id"]
        assert np.abs(r["sim"] - one["sim"]).max() < 1e-5 and np.abs(r["clips"] - one["clips"]).max() < 1e-5
        assert (r["counts"] == one["counts"][r["mine"]]).all()
        # descriptors come from the same kernels on another GPU of the same type: identical up to the last bit of the f32 descriptor
        full1 = one["dense"][r["mine"]] + one["dense_lo"][r["mine"]]
        assert np.abs((r["dense"] + r["dense_lo"]) - full1).max() < 1e-4
    assert (merged == one["ins"]).all() and len(one["objects"]) > 3


def _run_batch(rank, world, port, q):
    import torch.distributed as dist
    from ovo_b200 import synth
    from ovo_b200.map import SemanticMap
    from ovo_b200.p2p import VoteExchange
    from ovo_b200.sharding import ShardedBatchAssociation, shard_of_points
    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    K = synth.intrinsics(); F = 6
    xyz, ids, ins = synth.point_map(150000, synth.depth_map(frame_id=0), K, synth.pose(0), seed=9, frac_visible=0.6)
    frames = []
    for i in range(F):
        seg, bm = synth.grid_masks(rows=(6 if i % 2 == 0 else 3), cols=(8 if i % 2 == 0 else 5))
        frames.append((synth.depth_map(frame_id=3 * i), seg, synth.pose(3 * i), bm.shape[0]))
    dd = [torch.from_numpy(f[0]).to(dev) for f in frames]
    sd = [torch.from_numpy(f[1]).to(dev) for f in frames]
    c2ws, nms = [f[2] for f in frames], [f[3] for f in frames]
    sm_full, sm = SemanticMap(dev), SemanticMap(dev)
    xf, inf_ = torch.from_numpy(xyz).to(dev), torch.from_numpy(ins).to(dev)
    v_ref, nm_ref, nxt_ref = sm_full.associate_batch(xf, inf_, dd, sd, c2ws, K, 0, nms, track_th=60)
    mine = np.nonzero(shard_of_points(xyz, world) == rank)[0]
    out = {}
    for mode in ("nccl", "p2p", "p2p_persistent"):   # the last: the whole batch in one persistent launch (OVO_B200_VOTE)
        os.environ["OVO_B200_VOTE"] = "persistent" if mode == "p2p_persistent" else "launches"
        inf_ = torch.from_numpy(ins).to(dev)
        sm_full.associate_batch(xf, inf_, dd, sd, c2ws, K, 0, nms, track_th=60)
        xl, il = torch.from_numpy(xyz[mine]).to(dev), torch.from_numpy(ins[mine]).to(dev)
        if mode == "nccl":
            tables = torch.zeros(SemanticMap.batch_tables_size(0, nms), dtype=torch.int32, device=dev)
            v, nm, nxt = ShardedBatchAssociation(sm).associate(xl, il, dd, sd, c2ws, K, 0, nms, tables, track_th=60, n_frames=F)
        else:
            x = VoteExchange(sm, rank, world, dev, table_ints=4 + 64 * (64 * F + 1), slots=F)
            v, nm, nxt = x.associate_batch(xl, il, dd, sd, c2ws, K, 0, nms, track_th=60)
            v2, nm2, nxt2 = x.associate_batch(xl, il, dd, sd, c2ws, K, nxt, nms, track_th=60)   # a second batch re-uses the inbox (parities)
            r2 = sm_full.associate_batch(xf, inf_, dd, sd, c2ws, K, nxt_ref, nms, track_th=60)
            out["second_" + mode] = nxt2 == r2[2] and nm2 == r2[1] and all((v2[f][k] == r2[0][f][k]).all() for f in range(F) for k in v2[f]) \
                and bool(torch.equal(il, inf_[torch.from_numpy(mine).to(dev)]))
            inf_ = torch.from_numpy(ins).to(dev)
            sm_full.associate_batch(xf, inf_, dd, sd, c2ws, K, 0, nms, track_th=60)
        ok = nxt == nxt_ref and nm == nm_ref and all((v[f][k] == v_ref[f][k]).all() for f in range(F) for k in v[f])
        if mode == "nccl":
            ok = ok and bool(torch.equal(il, inf_[torch.from_numpy(mine).to(dev)]))
        out[mode] = bool(ok)
    torch.cuda.synchronize()
    q.put(dict(rank=rank, **out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_batch_association_nccl_and_fused_peer_exchange():
    for r in _spawn(_run_batch, 2):
        assert r["nccl"] and r["p2p"] and r["p2p_persistent"] and r["second_p2p"] and r["second_p2p_persistent"], r
