"""Pins the SAM-2 restatement (oracle/sam.py) to the reference's own outputs in tests/golden/sam_tiny.npz
(SAM2Base / SAM2ImagePredictor / SAM2AutomaticMaskGenerator + OVO's MaskGenerator.segment, run unmodified by
oracle/gen_golden.py gen_sam with the tiny Hiera geometry and our seeded weights).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden as GG, masks as OM, sam as OS
from ovo_b200.sam_config import SamConfig, random_state_dict, tiny_sam_config


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "sam_tiny.npz"))


@pytest.fixture(scope="module")
def tiny():
    cfg = tiny_sam_config()
    return cfg, random_state_dict(cfg, seed=0)


@pytest.fixture(scope="module")
def features(tiny):
    cfg, sd = tiny
    with torch.no_grad():
        px = OS.preprocess(GG.sam_image(), cfg.image_size)
        return px, OS.forward_image(px, sd, cfg)


def test_block_geometry_of_hiera_l():
    b = SamConfig().blocks()
    assert len(b) == 48 and [x.window for x in b].count(0) == 3 and all(b[i].window == 0 for i in (23, 33, 43))
    assert [i for i, x in enumerate(b) if x.q_pool] == [2, 8, 44]
    assert (b[2].dim, b[2].dim_out, b[2].heads, b[2].window) == (144, 288, 4, 8)      # window lags one block (hieradet.py:222-231)
    assert all(x.dim_out // x.heads == 72 for x in b)
    assert SamConfig().channel_list() == [1152, 576, 288, 144]


def test_transform_and_image_features_match_reference(features, gold):
    px, (emb, s0, s1) = features
    assert np.abs(px[0, :, ::16, ::16].numpy() - gold["px_sub"]).max() < 1e-5
    assert np.abs(emb[0, :, ::4, ::4].numpy() - gold["embed_sub"]).max() < 1e-4
    assert np.abs(s0[0, :, ::16, ::16].numpy() - gold["s0_sub"]).max() < 1e-4
    assert np.abs(s1[0, :, ::8, ::8].numpy() - gold["s1_sub"]).max() < 1e-4


def test_prompted_masks_match_reference(tiny, features, gold):
    cfg, sd = tiny
    _, (emb, s0, s1) = features
    pts = torch.from_numpy(OS.amg_points(16, 480, 640, cfg.image_size))
    with torch.no_grad():
        low, iou = OS.predict(pts[:64], emb, s0, s1, sd, cfg)
    assert np.abs(iou.numpy() - gold["iou64"]).max() < 1e-5
    low = low.clamp(-32, 32)
    assert np.abs(low[:, :, ::16, ::16].numpy() - gold["low_sub"]).max() < 2e-4
    assert np.abs(low[0, 0].numpy() - gold["low_full1"].astype(np.float32)).max() < 2e-2     # stored as f16


def test_generate_matches_reference(tiny, gold):
    cfg, sd = tiny
    img = GG.sam_image(*GG.SAM_AMG_HW, seed=6)
    r = OS.generate(img, sd, cfg, points_per_side=16, pred_iou_thresh=GG.SAM_THR["pred_iou_thresh"],
                    stability_thresh=GG.SAM_THR["stability_score_thresh"], box_nms_thresh=GG.SAM_THR["box_nms_thresh"])
    n = int(gold["n"])
    assert len(r["iou"]) == n
    assert np.abs(r["iou"] - gold["pred_iou"]).max() < 1e-5
    assert np.abs(r["stability"] - gold["stability"]).max() < 1e-3
    area = r["masks"].reshape(n, -1).sum(1)
    # the oracle's logits differ from the reference's by ~1e-5: a handful of pixels sitting on the threshold may flip
    assert np.abs(area - gold["area"]).max() <= 3
    H, W = GG.SAM_AMG_HW
    ref8 = np.unpackbits(gold["seg_bits_every8"])[: ((n + 7) // 8) * H * W].reshape(-1, H, W).astype(bool)
    assert (ref8 != r["masks"][::8]).reshape(len(ref8), -1).sum(1).max() <= 3
    xywh = r["boxes"].astype(np.float32).copy()
    xywh[:, 2:] -= xywh[:, :2]
    assert np.abs(xywh - gold["bbox"]).max() <= 1
    # points the masks came from (automatic_mask_generator.py:320-324)
    pts = OS.point_grid(16) * np.array([[W, H]])
    assert np.abs(pts[r["src"] // 3] - gold["points"]).max() < 1e-3
    # OVO's second stage (MaskGenerator.segment, mask_generator.py:113-119) on the oracle's proposals
    keep = OM.masks_update(r["masks"], r["iou"], r["stability"], 0.8, GG.SAM_OVO_SCORE_THR, 0.5)
    assert len(keep) == int(gold["ovo_n"])
    seg, bm, _ = OM.mask2segmap(r["masks"][keep], r["stability"][keep])
    assert (seg != gold["ovo_seg_map"]).sum() <= 6


def test_box_nms_matches_torchvision():
    """torchvision.ops.batched_nms is an un-vendored dependency of the reference (automatic_mask_generator.py:279);
    the installed torchvision is the oracle's oracle."""
    from torchvision.ops import batched_nms
    rng = np.random.default_rng(0)
    for n in (1, 7, 200):
        xy = rng.integers(0, 500, (n, 2)); wh = rng.integers(1, 200, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.uniform(0, 1, n).astype(np.float32)
        scores[n // 2:] = scores[: n - n // 2]          # ties
        ref = batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.zeros(n), 0.7).numpy()
        assert (OS.box_nms(boxes, scores, 0.7) == ref).all()
