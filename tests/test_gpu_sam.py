"""GPU parity of the SAM-2 mask proposal against the CPU oracle (oracle/sam.py, pinned to the reference by
tests/test_oracle_sam.py) and the reference's golden vectors (tests/golden/sam_tiny.npz), all through the C ABI.

Floating-point outputs (features, mask logits, predicted IoU): relative-L2 <= 2e-2 and 1 - cos <= 1e-3 against the f32
oracle (bf16 operands, f32 accumulation — the reference itself runs this path under bf16 autocast, mask_generator.py:44-46,112).
Integer outputs (the AMG filters, boxes, box NMS, final seg-map) are bit-exact given the same logits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import gen_golden as GG, masks as OM, sam as OS
from ovo_b200.sam_config import SamConfig, random_state_dict, tiny_sam_config

REL_TOL = 2e-2
COS_TOL = 1e-3


def _check(out, ref, what, rel_tol=REL_TOL, cos_tol=COS_TOL):
    out, ref = out.float().cpu(), torch.as_tensor(ref).float()
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    assert not torch.isnan(out).any(), what
    rel = ((out - ref).norm() / ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(out.reshape(-1, out.shape[-1]), ref.reshape(-1, ref.shape[-1]), dim=-1)
    assert rel < rel_tol, (what, rel)
    assert (1 - cos).max().item() < cos_tol, (what, (1 - cos).max().item())
    return rel


def _tok(x):
    """reference NCHW feature [1,C,H,W] -> token-major [H*W, C]"""
    return x[0].flatten(1).T.contiguous()


@pytest.fixture(scope="module")
def tiny():
    from ovo_b200.sam import Sam2
    cfg = tiny_sam_config()
    sd = random_state_dict(cfg, seed=0)
    return Sam2(cfg, sd, max_h=480, max_w=640, max_prompts=256), cfg, sd


@pytest.fixture(scope="module")
def oracle_features(tiny):
    _, cfg, sd = tiny
    with torch.no_grad():
        px = OS.preprocess(GG.sam_image(), cfg.image_size)
        taps = {}
        emb, s0, s1 = OS.forward_image(px, sd, cfg, taps)
    return px, emb, s0, s1, taps


def test_trunk_block_by_block(tiny, oracle_features):
    """Hiera trunk (hieradet.py:274-291): patch embed + pos embed, windowed blocks, q-pool transitions, global block."""
    sam, cfg, sd = tiny
    px, _, _, _, taps = oracle_features
    ref0 = taps["patch"][0].reshape(-1, cfg.embed_dim)
    _check(sam.set_pixels(px[0].cuda(), n_blocks=0), ref0, "patch embed")
    for i in range(len(cfg.blocks())):
        ref = taps[f"block{i}"][0]
        _check(sam.set_pixels(px[0].cuda(), n_blocks=i + 1), ref.reshape(-1, ref.shape[-1]), f"block {i}")


def test_image_features(tiny, oracle_features, golden_dir):
    """transform + trunk + neck + conv_s0/s1 + no_mem_embed (sam2_image_predictor.py:86-127) vs oracle and reference."""
    sam, cfg, sd = tiny
    px_ref, emb_ref, s0_ref, s1_ref, _ = oracle_features
    px, emb, s0, s1 = sam.set_image(torch.from_numpy(GG.sam_image()).cuda(), taps=True)
    assert (px.cpu() - px_ref[0]).abs().max().item() < 1e-4
    _check(emb, _tok(emb_ref), "image_embed")
    _check(s0, _tok(s0_ref), "feat_s0")
    _check(s1, _tok(s1_ref), "feat_s1")
    gold = np.load(os.path.join(golden_dir, "sam_tiny.npz"))
    g = sam.g
    e = emb.cpu().T.reshape(256, g, g)[:, ::4, ::4]
    assert (e - torch.from_numpy(gold["embed_sub"])).norm() / np.linalg.norm(gold["embed_sub"]) < REL_TOL
    assert np.abs(px.cpu().numpy()[:, ::16, ::16] - gold["px_sub"]).max() < 1e-4


def test_prompted_masks(tiny, oracle_features, golden_dir):
    """prompt encoder + two-way transformer + upscaling + hypernetworks + IoU head (sam2_image_predictor.py:337-432)."""
    sam, cfg, sd = tiny
    _, emb_ref, s0_ref, s1_ref, _ = oracle_features
    sam.set_image(torch.from_numpy(GG.sam_image()).cuda())
    pts = torch.from_numpy(OS.amg_points(16, 480, 640, cfg.image_size))
    low, iou = sam.predict(pts[:64].cuda())
    with torch.no_grad():
        low_ref, iou_ref = OS.predict(pts[:64], emb_ref, s0_ref, s1_ref, sd, cfg)
    assert (iou.cpu() - iou_ref).abs().max().item() < 1e-2
    _check(low.reshape(64 * 3, -1), low_ref.reshape(64 * 3, -1), "low_res_masks", rel_tol=3e-2)
    gold = np.load(os.path.join(golden_dir, "sam_tiny.npz"))
    assert np.abs(iou.cpu().numpy() - gold["iou64"]).max() < 1e-2
    sub = low.clamp(-32, 32)[:, :, ::16, ::16].cpu().numpy()
    assert np.linalg.norm(sub - gold["low_sub"]) / np.linalg.norm(gold["low_sub"]) < 3e-2
    # sign agreement of the logits (what the masks are made of)
    agree = ((low.cpu() > 0) == (low_ref > 0)).float().mean().item()
    assert agree > 0.99, agree


def test_amg_postprocess_bit_exact(tiny):
    """Filters, stability score, boxes and box NMS on GIVEN logits (automatic_mask_generator.py:294-375, utils/amg.py)."""
    sam, cfg, sd = tiny
    sam.set_image(torch.from_numpy(GG.sam_image()).cuda())
    pts = torch.from_numpy(OS.amg_points(16, 480, 640, cfg.image_size))
    low, iou = sam.predict(pts.cuda())
    for (H, W, a, b, nms) in ((480, 640, 0.45, 0.4, 1.0), (240, 320, 0.5, 0.5, 0.9999), (96, 128, 0.4, 0.3, 0.7)):
        prm = sam.amg_params(pred_iou_thresh=a, stability_score_thresh=b, box_nms_thresh=nms)
        out = sam.postprocess(low, iou, H, W, prm)
        ref = OS.amg_postprocess(low.cpu(), iou.cpu(), H, W, a, b, 1.0, nms)
        K = len(ref["iou"])
        assert out["masks"].shape[0] == K, (out["masks"].shape, K)
        if K == 0:
            continue
        assert (out["src"].cpu().numpy() == ref["src"]).all()
        assert (out["iou"].cpu().numpy() == ref["iou"]).all()
        assert (out["stability"].cpu().numpy() == ref["stability"]).all()
        assert (out["boxes"].cpu().numpy() == ref["boxes"]).all()
        assert (out["masks"].cpu().numpy().astype(bool) == ref["masks"]).all()


def test_box_nms_suppression_matches_oracle(tiny):
    """Localised synthetic logits (random weights give image-wide blobs, so the box NMS above rarely suppresses)."""
    sam, cfg, sd = tiny
    rng = np.random.default_rng(0)
    P, h = 40, 4 * sam.g
    low = np.full((P, 3, h, h), -5.0, np.float32)
    for p in range(P):
        for m in range(3):
            y, x = rng.integers(0, h - 60, 2)
            hh, ww = rng.integers(20, 60, 2)
            if m == 2 and p > 0:                      # near duplicate of the previous prompt's box
                y, x, hh, ww = last
            low[p, m, y:y + hh, x:x + ww] = 5.0 + rng.normal(0, 2.0, (hh, ww))
            last = (y, x + 1, hh, ww)
    iou = rng.uniform(0.3, 1.0, (P, 3)).astype(np.float32)
    iou[5] = iou[4]                                   # score ties
    prm = sam.amg_params(pred_iou_thresh=0.5, stability_score_thresh=0.5, box_nms_thresh=0.7)
    out = sam.postprocess(torch.from_numpy(low).cuda(), torch.from_numpy(iou).cuda(), 300, 400, prm)
    ref = OS.amg_postprocess(torch.from_numpy(low), torch.from_numpy(iou), 300, 400, 0.5, 0.5, 1.0, 0.7)
    assert 0 < len(ref["iou"]) < (iou > 0.5).sum()
    assert (out["src"].cpu().numpy() == ref["src"]).all()
    assert (out["boxes"].cpu().numpy() == ref["boxes"]).all()
    assert (out["masks"].cpu().numpy().astype(bool) == ref["masks"]).all()


def test_generate_end_to_end(tiny):
    """MaskGenerator.segment (mask_generator.py:102-120): the fused on-device pipeline equals the oracle's post-processing
    (AMG filters + OVO's masks_update + mask2segmap) applied to the same logits."""
    sam, cfg, sd = tiny
    H, W = GG.SAM_AMG_HW
    img = torch.from_numpy(GG.sam_image(H, W, seed=6)).cuda()
    prm = sam.amg_params(pred_iou_thresh=0.5, stability_score_thresh=0.5, box_nms_thresh=1.0, nms_score_th=GG.SAM_OVO_SCORE_THR)
    seg, maps = sam.generate(img, prm)
    pts = torch.from_numpy(OS.amg_points(16, H, W, cfg.image_size))
    low, iou = sam.predict(pts.cuda())                # same image is still set
    ref = OS.amg_postprocess(low.cpu(), iou.cpu(), H, W, 0.5, 0.5, 1.0, 1.0)
    keep = OM.masks_update(ref["masks"], ref["iou"], ref["stability"], 0.8, GG.SAM_OVO_SCORE_THR, 0.5)
    assert maps.shape[0] == len(keep) and len(keep) > 0
    seg_ref, bm_ref, _ = OM.mask2segmap(ref["masks"][keep], ref["stability"][keep])
    assert (seg.cpu().numpy() == seg_ref).all()
    assert (maps.cpu().numpy() == bm_ref).all()


def test_hiera_l_full_size():
    """SAM-2.1 Hiera-L geometry (sam2.1_hiera_l.yaml): 48 blocks, 212 M parameters, 1024^2 input."""
    from ovo_b200.sam import Sam2
    cfg = SamConfig()
    sd = random_state_dict(cfg, seed=1)
    sam = Sam2(cfg, sd, max_h=480, max_w=640, max_prompts=64)
    img = GG.sam_image(seed=7)
    with torch.no_grad():
        px = OS.preprocess(img, cfg.image_size)
        emb_ref, s0_ref, s1_ref = OS.forward_image(px, sd, cfg)
    px_g, emb, s0, s1 = sam.set_image(torch.from_numpy(img).cuda(), taps=True)
    assert (px_g.cpu() - px[0]).abs().max().item() < 1e-4
    _check(emb, _tok(emb_ref), "image_embed (Hiera-L)", rel_tol=3e-2)
    _check(s0, _tok(s0_ref), "feat_s0 (Hiera-L)")
    _check(s1, _tok(s1_ref), "feat_s1 (Hiera-L)")
    pts = torch.from_numpy(OS.amg_points(16, 480, 640, cfg.image_size))[::4][:16]
    low, iou = sam.predict(pts.cuda())
    with torch.no_grad():
        low_ref, iou_ref = OS.predict(pts, emb_ref, s0_ref, s1_ref, sd, cfg)
    assert (iou.cpu() - iou_ref).abs().max().item() < 2e-2
    _check(low.reshape(16 * 3, -1), low_ref.reshape(16 * 3, -1), "low_res_masks (Hiera-L)", rel_tol=5e-2, cos_tol=2e-3)


def test_batched_frames_equal_single_frames():
    """One trunk pass over several frames (ovo_sam_set_images / ovo_sam_generate_batch) gives every frame exactly what a
    single-frame call gives it: batching only folds the frames into the row dimension of the GEMMs / the window index."""
    from ovo_b200.sam import Sam2
    cfg = tiny_sam_config()
    sd = random_state_dict(cfg, seed=0)
    sam = Sam2(cfg, sd, max_h=240, max_w=320, max_prompts=256, max_batch=3)
    H, W = GG.SAM_AMG_HW
    imgs = torch.from_numpy(np.stack([GG.sam_image(H, W, seed=20 + i) for i in range(3)])).cuda()
    prm = sam.amg_params(pred_iou_thresh=0.5, stability_score_thresh=0.5, box_nms_thresh=1.0, nms_score_th=GG.SAM_OVO_SCORE_THR)
    pts = torch.from_numpy(OS.amg_points(16, H, W, cfg.image_size))[:32].cuda()
    single = []
    for i in range(3):
        _, emb, s0, s1 = sam.set_image(imgs[i], taps=True)
        low, iou = sam.predict(pts)
        single.append((emb.clone(), low.clone(), iou.clone(), sam.generate(imgs[i], prm)))
    sam.set_images(imgs)
    for i in range(3):
        sam.select_image(i)
        low, iou = sam.predict(pts)
        assert torch.equal(low, single[i][1]) and torch.equal(iou, single[i][2]), f"frame {i}"
    out = sam.generate_batch(imgs, prm)
    for i in range(3):
        assert torch.equal(out[i][0], single[i][3][0]) and torch.equal(out[i][1], single[i][3][1]), f"frame {i}"
    assert len({int(o[1].shape[0]) for o in out}) >= 1


def test_edge_cases_and_errors(tiny):
    """No surviving mask (automatic_mask_generator.py:331-342 filters everything; OVO returns None upstream, ovo.py:141-143),
    frames larger than the handle was created for, too many prompts: empty results / loud errors, never garbage."""
    sam, cfg, sd = tiny
    img = torch.from_numpy(GG.sam_image(seed=12)).cuda()
    seg, maps = sam.generate(img, sam.amg_params(pred_iou_thresh=0.999, stability_score_thresh=0.999))
    assert maps.shape[0] == 0 and int(seg.max()) == -1
    low, iou = sam.predict(torch.from_numpy(OS.amg_points(16, 480, 640, cfg.image_size))[:8].cuda())
    out = sam.postprocess(low, iou, 480, 640, sam.amg_params(pred_iou_thresh=0.999))
    assert out["masks"].shape[0] == 0
    with pytest.raises(RuntimeError, match="exceeds"):
        sam.set_image(torch.zeros(481, 640, 3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError, match="prompts"):
        sam.predict(torch.zeros(sam.max_prompts + 1, 2, device="cuda"))
    with pytest.raises(RuntimeError, match="frames"):
        sam.set_images(torch.zeros(2, 480, 640, 3, dtype=torch.uint8, device="cuda"))     # handle created with max_batch 1
    # the handle still works after the errors
    seg2, maps2 = sam.generate(img, sam.amg_params(pred_iou_thresh=0.45, stability_score_thresh=0.4, box_nms_thresh=0.9999, nms_score_th=0.2))
    assert maps2.shape[0] > 0
