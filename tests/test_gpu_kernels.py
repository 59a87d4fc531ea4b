"""GPU parity of the map / query kernels against the CPU oracle and the reference's golden vectors.
Everything goes through the C ABI (ovo_b200._lib -> libovo_b200.so)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fusion as OF, gen_golden as GG
from ovo_b200 import synth


@pytest.fixture(scope="module")
def sm():
    from ovo_b200.map import SemanticMap
    return SemanticMap()


def _dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 192), (1154, 3072, 1024), (77, 1024, 4096), (5000, 20, 1024)])
def test_gemm_matches_f32(M, N, K):
    """tcgen05 GEMM (every N-tile width) vs an f32 matmul of the same bf16 operands; tolerance 1e-5 relative-L2."""
    from ovo_b200.encoder import gemm_bf16
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().T + bias
    for bn in (0, 32, 64, 128, 256):
        out = gemm_bf16(A, B, bias, force_bn=bn)
        assert ((out - ref).norm() / ref.norm()).item() < 1e-5, bn


@pytest.mark.parametrize("N,Q", [(1, 1), (129, 20), (70000, 21), (20000, 200), (3000, 300)])
def test_query_dense(sm, N, Q):
    """Dense cosine query == f32 matmul of the bf16 bank with the bf16-rounded text bank (tolerance 1e-5 rel-L2),
    and within 1e-3 absolute of the unrounded f32 text bank."""
    torch.manual_seed(1)
    bank = torch.nn.functional.normalize(torch.randn(N, 1024, device="cuda"), dim=-1).bfloat16()
    text = torch.nn.functional.normalize(torch.randn(Q, 1024, device="cuda"), dim=-1)
    out = sm.query_dense(bank, text)
    ref = bank.float() @ text.bfloat16().float().T
    assert ((out - ref).norm() / ref.norm()).item() < 1e-5
    assert (out - bank.float() @ text.T).abs().max().item() < 1e-3
    cls, conf = sm.classify(out, 0.0)
    mx, am = out.max(1)
    assert (cls.long() == torch.where(mx > 0, am, -1)).all()
    assert torch.equal(conf, torch.where(mx > 0, mx, torch.zeros_like(mx)))


def test_query_linearity_full_size(sm):
    """Size-independent property at BASELINE size (2M points, Q=20): the query is linear in the text bank."""
    N = 2_000_000
    g = torch.Generator(device="cuda").manual_seed(2)
    bank = torch.randn(N, 1024, device="cuda", generator=g).bfloat16()
    t1 = torch.randn(20, 1024, device="cuda", generator=g).bfloat16().float()
    t2 = torch.randn(20, 1024, device="cuda", generator=g).bfloat16().float()
    a, b = sm.query_dense(bank, t1), sm.query_dense(bank, t2)
    c = sm.query_dense(bank, (t1 + t2) / 2)        # exactly representable? no -> compare with tolerance
    ref = (a + b) / 2
    assert ((c - ref).norm() / ref.norm()).item() < 5e-3
    idx = torch.randint(0, N, (4096,), device="cuda")
    assert ((a[idx] - bank[idx].float() @ t1.T).abs().max().item()) < 2e-2 * 32   # spot check against f32


def test_query_config4_size_5m_points_200_classes(sm):
    """BASELINE config 4's query: 5M-point dense map (10.2 GB bf16) against a 200-class text bank; spot-checked against f32 and
    classified (argmax + threshold, ovo.py:486-491) on the device."""
    N, Q = 5_000_000, 200
    g = torch.Generator(device="cuda").manual_seed(4)
    bank = torch.nn.functional.normalize(torch.randn(N, 1024, device="cuda", generator=g), dim=-1).bfloat16()
    text = torch.nn.functional.normalize(torch.randn(Q, 1024, device="cuda", generator=g), dim=-1)
    out = sm.query_dense(bank, text)
    assert out.shape == (N, Q)
    idx = torch.randint(0, N, (8192,), device="cuda", generator=g)
    ref = bank[idx].float() @ text.bfloat16().float().T
    assert (out[idx] - ref).abs().max().item() < 1e-4
    cls, conf = sm.classify(out, 0.05)
    mx, am = out.max(1)
    assert (cls.long() == torch.where(mx > 0.05, am, -1)).all()
    del bank, out


def test_query_instances_and_fuse_views(sm):
    torch.manual_seed(3)
    store = torch.randn(40, 256, device="cuda")
    bank = torch.zeros(8, 256, device="cuda")
    idx = torch.tensor([0, 5, 9, 3, 7, 7, 8, 2, 1], dtype=torch.int32, device="cuda")
    off = torch.tensor([0, 3, 4, 9], dtype=torch.int32, device="cuda")
    rows = torch.tensor([2, 0, 5], dtype=torch.int32, device="cuda")
    for mode in (0, 1, 2):
        chosen = torch.zeros(3, dtype=torch.int32, device="cuda")
        sm.fuse_views(store, idx, off, mode, bank, rows, chosen)
        for j in range(3):
            clips = store[idx[off[j]:off[j + 1]].long()]
            if mode == 0 or clips.shape[0] == 1:
                ref = clips.mean(0)
            elif mode == 1:
                ref = clips[torch.abs(clips[None] - clips[:, None]).sum((1, 2)).argmin()]
            else:
                ref = clips[torch.cosine_similarity(clips[None], clips[:, None], dim=-1).sum(-1).argmax()]
            assert (bank[rows[j]] - ref).abs().max().item() < 1e-5, (mode, j)
    text = torch.randn(5, 256, device="cuda")
    out = sm.query_instances(bank, text, rows)
    assert (out - bank[rows.long()] @ text.T).abs().max().item() < 1e-4


def test_merge_masks(sm):
    rng = np.random.default_rng(0)
    masks = (rng.random((6, 48, 64)) > 0.7)
    group = np.array([0, 1, 0, -1, 2, 1], np.int32)
    out, areas = sm.merge_masks(*_dev(masks.astype(np.uint8), group), 3)
    for r in range(3):
        ref = np.any(masks[group == r], axis=0)
        assert (out[r].cpu().numpy().astype(bool) == ref).all() and int(areas[r]) == ref.sum()


@pytest.mark.parametrize("fid,N,fv", GG.ASSOC_CASES)
def test_association_matches_reference_golden(sm, golden_dir, fid, N, fv):
    """CUDA association == the reference's own geometry functions (golden) bit for bit."""
    g = np.load(os.path.join(golden_dir, "assoc.npz"))
    K = synth.intrinsics(); c2w = synth.pose(fid); d = synth.depth_map(frame_id=fid)
    seg, _ = synth.grid_masks()
    xyz, ids, ins = synth.point_map(N, d, K, c2w, seed=fid, frac_visible=fv)
    xyz_d, ins_d, d_d, seg_d = _dev(xyz, ins, d, seg)
    assert (np.packbits(sm.depth_filter(d_d).cpu().numpy() == -1) == g[f"depth_rejected_{fid}"]).all()
    votes, n_matched, nxt = sm.associate(xyz_d, ins_d, d_d, seg_d, c2w, K, 0, kf_slot=1)
    ref_seg = g[f"seg_of_pt_{fid}"].astype(np.int32)
    assert n_matched == (ref_seg > -2).sum()
    pairs = sm.matches(1, int(votes["n_matched"].sum())).cpu().numpy()
    mine = np.full(N, -2, np.int32)
    mine[ref_seg == -1] = -1                         # matched but outside every mask: not listed
    mine[pairs[:, 0]] = pairs[:, 1]
    assert (mine == ref_seg).all()
    new, rows, nxt_o = OF.track(ins, ref_seg, seg, 100, 0)
    assert nxt == nxt_o and (ins_d.cpu().numpy() == new).all()
    for k in votes:
        assert (votes[k] == np.array([r[k] for r in rows])).all(), k


@pytest.mark.parametrize("N,track_th,ratio", [(0, 100, ()), (1, 0, ()), (257, 3, ()), (150000, 100, ()),
                                              (90000, 50, (1.0, 1.0, 0)), (60000, 100, (2.0, 2.0, 4))])
def test_association_multi_keyframe_vs_oracle(sm, N, track_th, ratio):
    """Four keyframes with changing masks/poses: votes, new ids, ties and the dense running mean are
    bit-exact against oracle/fusion.py, including empty / tiny maps and the rgb-depth ratio fix-up."""
    K = synth.intrinsics()
    h, w = 480, 640
    H, W = (h, w) if not ratio else (int((h + 2 * ratio[2]) * ratio[0]), int((w + 2 * ratio[2]) * ratio[1]))
    d0 = synth.depth_map(frame_id=0)
    xyz, ids, ins = synth.point_map(max(N, 1), d0, K, synth.pose(0), seed=5, frac_visible=0.6)
    xyz, ins = xyz[:N], ins[:N]
    xyz_d, ins_d = _dev(xyz, ins)
    D = 64
    bank = torch.zeros(max(N, 1), D, device="cuda", dtype=torch.bfloat16)
    bank_lo = torch.zeros_like(bank)
    counts = torch.zeros(max(N, 1), device="cuda", dtype=torch.int32)
    hi_o = np.zeros((max(N, 1), D), np.float32); lo_o = np.zeros_like(hi_o); counts_o = np.zeros(max(N, 1), np.int32)
    next_id = 0
    for i in range(4):
        fid = 4 * i
        c2w = synth.pose(fid); d = synth.depth_map(frame_id=fid)
        seg, bm = synth.grid_masks(H, W, rows=(6 if i % 2 == 0 else 3), cols=(8 if i % 2 == 0 else 5))
        d_d, seg_d = _dev(d, seg)
        votes, nm, nxt = sm.associate(xyz_d, ins_d, d_d, seg_d, c2w, K, next_id, track_th=track_th,
                                      rgb_depth_ratio=ratio, kf_slot=i)
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        seg_of_pt, _ = OF.associate(xyz, ins, d, seg, c2w, w2c, K, 0.05, True, ratio) if N else (np.zeros(0, np.int32), None)
        new, rows, nxt_o = OF.track(ins, seg_of_pt, seg, track_th, next_id)
        assert nm == (seg_of_pt > -2).sum() and nxt == nxt_o
        assert (ins_d.cpu().numpy() == new).all()
        for k in votes:
            assert (votes[k] == np.array([r[k] for r in rows])).all(), (i, k)
        order, fused, mask_row = OF.fuse_masks(bm, rows)
        feats = torch.randn(max(len(order), 1), D, generator=torch.Generator().manual_seed(i))
        sm.fuse_dense(i, bank, bank_lo, counts, feats.cuda(), torch.from_numpy(mask_row).cuda())
        if N:
            OF.dense_fuse(hi_o, lo_o, counts_o, [np.where(seg_of_pt >= 0, seg_of_pt, -1)], [mask_row], feats.numpy())
        assert (counts.cpu().numpy() == counts_o).all()
        assert (bank.float().cpu().numpy() == hi_o).all() and (bank_lo.float().cpu().numpy() == lo_o).all()
        ins, next_id = new, nxt_o


def test_association_idempotent_at_full_size(sm):
    """BASELINE-size property (2M points): a second pass over the same keyframe creates no instance, moves no
    point, and every mask now votes for the instance it created."""
    K = synth.intrinsics(); d = synth.depth_map(); N = 2_000_000
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=0)
    seg, _ = synth.grid_masks()
    xyz_d, ins_d, d_d, seg_d = _dev(xyz, ins, d, seg)
    v1, n1, nxt1 = sm.associate(xyz_d, ins_d, d_d, seg_d, synth.pose(0), K, 0, kf_slot=0)
    after1 = ins_d.clone()
    v2, n2, nxt2 = sm.associate(xyz_d, ins_d, d_d, seg_d, synth.pose(0), K, nxt1, kf_slot=1)
    assert n1 == n2 and nxt2 == nxt1 and torch.equal(after1, ins_d)
    assert (v2["n_unassigned"] == 0).all() and (v2["is_new"] == 0).all()
    assert (v2["ins_id"] == v1["ins_id"]).all() and (v2["n_assigned"] == v1["n_matched"]).all()
    assert int((ins_d >= 0).sum()) == int(v1["n_matched"][v1["ins_id"] >= 0].sum())


def test_reserved_workspaces_give_the_same_association():
    """ovo_map_reserve only sizes workspaces: same votes / ids / match lists as the grow-on-demand handle."""
    from ovo_b200.map import SemanticMap
    K = synth.intrinsics(); d = synth.depth_map(); N = 300_000
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=4, frac_visible=0.5)
    seg, _ = synth.grid_masks()
    outs = []
    for reserve in (False, True):
        m = SemanticMap("cuda:0")
        if reserve:
            m.reserve(points=1_000_000, instances=5000, masks=300, matches=480 * 640)
        xyz_d, ins_d, d_d, seg_d = _dev(xyz, ins, d, seg)
        v, n, nxt = m.associate(xyz_d, ins_d, d_d, seg_d, synth.pose(0), K, 0, kf_slot=3)
        pairs = m.matches(3, n).cpu().numpy()
        outs.append((v, n, nxt, ins_d.cpu().numpy(), pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]))
    a, b = outs
    assert a[1] == b[1] and a[2] == b[2] and (a[3] == b[3]).all() and (a[4] == b[4]).all()
    for k in a[0]:
        assert (np.asarray(a[0][k]) == np.asarray(b[0][k])).all(), k
    with pytest.raises(RuntimeError):
        SemanticMap("cuda:0").reserve(masks=1 << 20, instances=1 << 20)


def _batch_scene(N, F, seed=9):
    K = synth.intrinsics()
    d0 = synth.depth_map(frame_id=0)
    xyz, ids, ins = synth.point_map(N, d0, K, synth.pose(0), seed=seed, frac_visible=0.6)
    frames = []
    for i in range(F):
        seg, bm = synth.grid_masks(rows=(6 if i % 2 == 0 else 3), cols=(8 if i % 2 == 0 else 5))
        frames.append(dict(depth=synth.depth_map(frame_id=3 * i), seg=seg, bm=bm, c2w=synth.pose(3 * i)))
    return K, xyz, ins, frames


@pytest.mark.parametrize("vote_mode", ["launches", "persistent"])
@pytest.mark.parametrize("N,F,track_th", [(120000, 5, 100), (0, 2, 100), (300, 3, 2), (70000, 19, 60)])
def test_associate_batch_equals_sequential_and_oracle(sm, monkeypatch, N, F, track_th, vote_mode):
    """Both vote schedules of the batch (one launch per keyframe — the default — and the single persistent launch,
    OVO_B200_VOTE=persistent).  One pass over the map for F keyframes (id decisions on the device, one host sync) == F single-keyframe associations ==
    oracle/fusion.py: votes, n_matched, next_ins_id, the per-point ids and the mask -> instance table, bit for bit; 19 keyframes
    exercise the second shared-memory group of the pass (16 keyframes at a time)."""
    monkeypatch.setenv("OVO_B200_VOTE", vote_mode)
    K, xyz, ins, frames = _batch_scene(max(N, 1), F)
    xyz, ins = xyz[:N], ins[:N]
    xyz_d, ins_a = _dev(xyz, ins)
    ins_b = ins_a.clone()
    dd = [torch.from_numpy(f["depth"]).cuda() for f in frames]
    sd = [torch.from_numpy(f["seg"]).cuda() for f in frames]
    nms = [int(f["bm"].shape[0]) for f in frames]
    seq, nxt = [], 0
    for i in range(F):
        v, nm, nxt = sm.associate(xyz_d, ins_a, dd[i], sd[i], frames[i]["c2w"], K, nxt, track_th=track_th, kf_slot=i, n_masks=nms[i])
        seq.append((v, nm))
    mask_ins = torch.full((F, 64), -7, dtype=torch.int32, device="cuda")
    votes, nmb, nxt_b = sm.associate_batch(xyz_d, ins_b, dd, sd, [f["c2w"] for f in frames], K, 0, nms, track_th=track_th,
                                           kf_slots=list(range(F)), mask_ins_out=mask_ins)
    assert nxt_b == nxt and torch.equal(ins_a, ins_b)
    for i in range(F):
        assert nmb[i] == seq[i][1]
        for k in votes[i]:
            assert (votes[i][k] == seq[i][0][k]).all(), (i, k)
        got = mask_ins[i].cpu().numpy()
        assert (got[: nms[i]] == votes[i]["ins_id"]).all() and (got[nms[i]:] == -1).all()
    # and the oracle, keyframe by keyframe
    ins_o, nxt_o = ins.copy(), 0
    for i, f in enumerate(frames):
        w2c = torch.linalg.inv(torch.from_numpy(f["c2w"])).numpy()
        seg_of_pt, _ = OF.associate(xyz, ins_o, f["depth"], f["seg"], f["c2w"], w2c, K, 0.05, True) if N else (np.zeros(0, np.int32), None)
        ins_o, rows, nxt_o = OF.track(ins_o, seg_of_pt, f["seg"], track_th, nxt_o)
        for k in votes[i]:
            assert (votes[i][k] == np.array([r[k] for r in rows])).all(), (i, k)
    assert nxt_o == nxt_b and (ins_b.cpu().numpy() == ins_o).all()


def test_fuse_dense_batch_vs_oracle_and_sequential(sm):
    """One pass over the two-plane bank for several keyframes: bit-exact against oracle/fusion.py dense_fuse (descriptors of a
    point summed in f32, one mean update), from match lists and from the dense rows of a batched association alike; equal to
    one pass per keyframe up to the rounding of the 16-bit mean (2^-15 relative)."""
    N, D, F = 120000, 128, 5
    K, xyz, ins, frames = _batch_scene(N, F)
    xyz_d, ins_d = _dev(xyz, ins)
    dd = [torch.from_numpy(f["depth"]).cuda() for f in frames]
    sd = [torch.from_numpy(f["seg"]).cuda() for f in frames]
    feats_all, rows_all, nm = [], [], 48
    base, nxt = 0, 0
    ins_o = ins.copy()
    segs_o = []
    for i in range(F):
        votes, _, nxt = sm.associate(xyz_d, ins_d, dd[i], sd[i], frames[i]["c2w"], K, nxt, kf_slot=i)
        w2c = torch.linalg.inv(torch.from_numpy(frames[i]["c2w"])).numpy()
        seg_of_pt, _ = OF.associate(xyz, ins_o, frames[i]["depth"], frames[i]["seg"], frames[i]["c2w"], w2c, K, 0.05, True)
        segs_o.append(np.where(seg_of_pt >= 0, seg_of_pt, -1))
        n_m = len(votes["ins_id"])
        local = np.where(votes["ins_id"] >= 0, np.arange(n_m), -1).astype(np.int32)
        local[::7] = -1                                   # some masks produce no descriptor
        feats_all.append(torch.randn(n_m, D, generator=torch.Generator().manual_seed(i)))
        row = np.full(nm, -1, np.int32); row[:n_m] = np.where(local >= 0, local + base, -1)
        rows_all.append((local, row)); base += n_m
    feats = torch.cat(feats_all).cuda()
    mk = lambda: (torch.zeros(N, D, device="cuda", dtype=torch.bfloat16), torch.zeros(N, D, device="cuda", dtype=torch.bfloat16),
                  torch.zeros(N, device="cuda", dtype=torch.int32))
    hi_a, lo_a, cnt_a = mk()
    hi_b, lo_b, cnt_b = mk()
    hi_c, lo_c, cnt_c = mk()
    off = 0
    for i in range(F):
        n_m = feats_all[i].shape[0]
        sm.fuse_dense(i, hi_a, lo_a, cnt_a, feats[off:off + n_m].contiguous(), torch.from_numpy(rows_all[i][0]).cuda())
        off += n_m
    mr = torch.from_numpy(np.stack([r[1] for r in rows_all])).cuda()
    sm.fuse_dense_batch(list(range(F)), hi_b, lo_b, cnt_b, feats, mr)
    # the same keyframes through the batched association (dense rows instead of lists)
    ins_e = torch.from_numpy(ins).cuda()
    sm.associate_batch(xyz_d, ins_e, dd, sd, [f["c2w"] for f in frames], K, 0, [int(f["bm"].shape[0]) for f in frames], kf_slots=list(range(F)))
    sm.fuse_dense_batch(list(range(F)), hi_c, lo_c, cnt_c, feats, mr)
    assert torch.equal(ins_e, ins_d)
    hi_o = np.zeros((N, D), np.float32); lo_o = np.zeros_like(hi_o); cnt_o = np.zeros(N, np.int32)
    OF.dense_fuse(hi_o, lo_o, cnt_o, segs_o, [r[1] for r in rows_all], feats.cpu().numpy())
    for hi, lo, cnt in ((hi_b, lo_b, cnt_b), (hi_c, lo_c, cnt_c)):
        assert (cnt.cpu().numpy() == cnt_o).all() and int(cnt.max()) >= 3
        assert (hi.float().cpu().numpy() == hi_o).all() and (lo.float().cpu().numpy() == lo_o).all()
    assert torch.equal(cnt_a, cnt_b)
    fa, fb = hi_a.float() + lo_a.float(), hi_b.float() + lo_b.float()
    assert (fa - fb).abs().max().item() <= 2.0 ** -14 * fb.abs().max().item()


def test_dense_bank_keeps_the_mean_over_500_views(sm):
    """a bf16 running mean re-rounded at every update freezes once the increment (e - f)/c drops below half
    an ulp.  The two-plane bank must not: 500 updates of the same points against the f64 mean, 1 - cos <= 1e-6 (hi + lo) and
    <= 1e-4 for the bf16 query plane alone."""
    N, D, V = 512, 1024, 500
    K = synth.intrinsics(); d = synth.depth_map()
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=1, frac_visible=1.0)
    seg, bm = synth.grid_masks()
    xyz_d, ins_d, d_d, seg_d = _dev(xyz, ins, d, seg)
    votes, nm, nxt = sm.associate(xyz_d, ins_d, d_d, seg_d, synth.pose(0), K, 0, track_th=0, kf_slot=0)
    pairs = sm.matches(0, nm).cpu().numpy()
    assert len(pairs) > 100
    M = bm.shape[0]
    g = torch.Generator().manual_seed(0)
    base = torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=-1)
    hi = torch.zeros(N, D, device="cuda", dtype=torch.bfloat16); lo = torch.zeros_like(hi)
    cnt = torch.zeros(N, device="cuda", dtype=torch.int32)
    mask_row = torch.arange(M, dtype=torch.int32, device="cuda")
    acc = torch.zeros(M, D, dtype=torch.float64)
    for v in range(V):
        e = torch.nn.functional.normalize(0.8 * base + 0.6 * torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=-1), dim=-1)
        acc += e.double()
        sm.fuse_dense(0, hi, lo, cnt, e.cuda(), mask_row)
    pts, msk = pairs[:, 0], pairs[:, 1]
    assert (cnt.cpu().numpy()[pts] == V).all()
    ref = (acc / V)[msk]
    full = (hi.double() + lo.double()).cpu()[pts]
    cos = torch.nn.functional.cosine_similarity
    assert (1 - cos(full, ref, dim=-1)).max().item() <= 1e-6
    assert (1 - cos(hi.double().cpu()[pts], ref, dim=-1)).max().item() <= 1e-4
    assert ((full - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item() <= 3e-3     # descriptors enter rounded to bf16


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_mask_nms_and_segmap_match_oracle(sm, golden_dir, seed):
    """S2: GPU mask NMS + seg-map painting == oracle/masks.py (pinned to the reference in tests/golden/masks.npz)."""
    from oracle import masks as OM
    masks, iou, stab = OM.synth_masks(seed=seed)
    keep = sm.mask_nms(torch.from_numpy(masks), torch.from_numpy(stab * iou)).cpu().numpy()
    ref = OM.masks_update(masks, iou, stab)
    assert np.nonzero(keep)[0].tolist() == ref.tolist() and 0 < len(ref) < len(masks)
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    if f"kept_{seed}" in g:
        assert np.nonzero(keep)[0].tolist() == g[f"kept_{seed}"].tolist()
    seg, maps, order = sm.mask2segmap(torch.from_numpy(masks[ref]), torch.from_numpy(stab[ref]))
    seg_o, maps_o, order_o = OM.mask2segmap(masks[ref], stab[ref])
    assert (seg.cpu().numpy() == seg_o).all() and (maps.cpu().numpy() == maps_o).all() and order.cpu().tolist() == order_o.tolist()


def test_map_producer_matches_oracle(sm):
    """PointMapper (ovo_map_integrate) == oracle/mapper.py point for point (bit-exact f32), colours included, with
    geometric buffer growth."""
    from oracle import mapper as OMp, gen_golden as GG2
    from ovo_b200.mapper import PointMapper
    K = synth.intrinsics()
    pm = PointMapper({"device": "cuda", "mapping": {"k_pooling": 3, "reserve_points": 100000}}, torch.from_numpy(K), semmap=sm)
    xyz = np.zeros((0, 3), np.float32)
    for i, (fid, pf, yaw) in enumerate(GG2.MAPPER_FRAMES):
        d, c2w, img = synth.depth_map(frame_id=fid), synth.pose(pf, yaw=yaw), synth.rgb(seed=fid)
        n_new = pm.map([fid, img, d, c2w], torch.from_numpy(c2w))
        new, pix = OMp.integrate_frame(xyz, d, c2w, K)
        assert n_new == len(new)
        got = pm.pcd[len(xyz):].cpu().numpy()
        assert (got == new).all()
        assert (pm.pcd_colors[len(xyz):].cpu().numpy() == img[pix[:, 0], pix[:, 1]]).all()
        xyz = np.concatenate([xyz, new])
    pts, pids, obj = pm.get_map()
    assert pts.shape[0] == len(xyz) and (pids.reshape(-1).cpu().numpy() == np.arange(len(xyz))).all() and int(obj.max()) == -1
    assert pm.capacity >= len(xyz) > 100000          # grew past the initial reservation


def test_associate_launch_wait_equals_associate(sm):
    """ovo_map_associate in two halves (everything enqueued, then ONE host synchronisation whenever the caller needs the rows):
    the same rows, ids and match list as the one-call form; a second launch before the wait is refused."""
    K = synth.intrinsics(); d = synth.depth_map(); N = 200_000
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=2, frac_visible=0.5)
    seg, _ = synth.grid_masks()
    xyz_d, ins_a, d_d, seg_d = _dev(xyz, ins, d, seg)
    ins_b = ins_a.clone()
    va, na, xa = sm.associate(xyz_d, ins_a, d_d, seg_d, synth.pose(0), K, 0, kf_slot=5)
    pa = sm.matches(5, na).cpu().numpy()
    sm.associate_launch(xyz_d, ins_b, d_d, seg_d, synth.pose(0), K, 0, kf_slot=6)
    with pytest.raises(RuntimeError):
        sm.associate_launch(xyz_d, ins_b, d_d, seg_d, synth.pose(0), K, 0, kf_slot=7)
    host_work = sum(i * i for i in range(10000))                  # (the caller's own work overlaps the GPU here)
    vb, nb, xb = sm.associate_wait()
    pb = sm.matches(6, nb).cpu().numpy()
    assert host_work > 0 and na == nb and xa == xb and torch.equal(ins_a, ins_b)
    for k in va:
        assert (va[k] == vb[k]).all(), k
    assert (pa[np.lexsort((pa[:, 1], pa[:, 0]))] == pb[np.lexsort((pb[:, 1], pb[:, 0]))]).all()


@pytest.mark.parametrize("world,n,cap", [(8, 76800, 14464), (2, 5000, 4000), (5, 3000, 500), (16, 1000, 200)])
def test_route_pack_kernel_matches_the_torch_rule(world, n, cap):
    """ovo_route_pack (one kernel) == the torch formulation of route_new_points_fixed: shard of every point = shard_of_points,
    slot = its rank among this rank's points of that shard in creation order, sentinels elsewhere, surplus counted."""
    from ovo_b200 import _lib
    from ovo_b200.sharding import shard_of_points
    rng = np.random.default_rng(world)
    xyz = rng.uniform(-8, 8, (n, 3)).astype(np.float32)
    xyz[::97] = np.round(xyz[::97] * 4) / 4                      # points exactly on voxel faces
    ids = (np.arange(n) + 1000).astype(np.int32)
    xd, idd = torch.from_numpy(xyz).cuda(), torch.from_numpy(ids).cuda()
    rec = torch.empty(world * cap, 4, device="cuda")
    ovf = torch.zeros((), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().ovo_route_pack(_lib.ptr(xd), _lib.ptr(idd), n, world, 0.25, cap, 1.0e6, _lib.ptr(rec), _lib.ptr(ovf),
                                         _lib.stream_ptr()), "ovo_route_pack")
    got = rec.cpu().numpy()
    dst = shard_of_points(xyz, world)
    exp = np.full((world * cap, 4), 1.0e6, np.float32)
    exp[:, 3] = np.int32(-1).view(np.float32)
    surplus = 0
    for d in range(world):
        sel = np.nonzero(dst == d)[0]
        surplus += max(0, len(sel) - cap)
        sel = sel[:cap]
        exp[d * cap: d * cap + len(sel), :3] = xyz[sel]
        exp[d * cap: d * cap + len(sel), 3] = ids[sel].view(np.float32)
    assert (got.view(np.int32) == exp.view(np.int32)).all() and int(ovf) == surplus
    assert (surplus > 0) == (world == 5)
