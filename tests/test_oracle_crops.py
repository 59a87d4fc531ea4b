"""Pins oracle/crops.py (the crop-based descriptor path: segmap2segimg, encode_image, fuse_clips) to the outputs
of the unmodified reference stored in tests/golden/crops.npz (oracle/gen_golden.py:gen_crops).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import crops as OC, gen_golden as GG
from ovo_b200 import synth
from ovo_b200.encoder import random_state_dict
from test_oracle_golden import _ocfg


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "crops.npz"))


@pytest.fixture(scope="module")
def setup():
    cfg = GG.tiny_cfg()
    return _ocfg(cfg), random_state_dict(cfg, seed=0), synth.rgb(480, 640, seed=21), GG.crop_masks()


def test_boxes_match_reference(setup, gold):
    _, _, _, bm = setup
    assert (OC.mask_boxes_xywh(bm) == gold["boxes_xywh"]).all()
    empty = OC.mask_boxes_xywh(np.zeros((1, 8, 8), bool))
    assert (empty == 0).all()


def test_seg_images_match_reference(setup, gold):
    _, _, img, bm = setup
    imt = torch.from_numpy(img.transpose(2, 0, 1).copy())
    seg = OC.seg_images(bm, imt, True, 50, 336)
    # uint8 after a float resize + round: the restated weights differ from ATen's in the last float bit, so a value
    # sitting on .5 may land on the other side
    d = np.abs(seg[:, :, ::6, ::6].numpy().astype(int) - gold["segimg_sub"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert np.abs(seg.long().sum((2, 3)).numpy() - gold["segimg_sum"]).max() <= 120      # <= 1e-3 of the 113k pixels flip by one
    segv = OC.seg_images(bm, imt, False, 50, 224)
    d = np.abs(segv[:, :, ::4, ::4].numpy().astype(int) - gold["segimg_vanilla_sub"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_encode_image_matches_reference(setup, gold):
    ocfg, sd, img, _ = setup
    imt = torch.from_numpy(img.transpose(2, 0, 1).copy())
    with torch.no_grad():
        e = OC.encode_image(imt[None].float() / 255.0, sd, ocfg, GG.CROP_POOL_HEADS)
    ref = gold["encode_image_global"]
    assert np.abs(e.numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("et,res", GG.CROP_CASES)
def test_extract_clip_matches_reference(setup, gold, et, res):
    ocfg, sd, img, bm = setup
    with torch.no_grad():
        f = OC.extract_clip(img, bm, sd, ocfg, et, mask_res=res, pool_heads=GG.CROP_POOL_HEADS)
    ref = torch.from_numpy(gold[f"{et}_{res}"])
    assert f.shape == ref.shape
    cos = torch.nn.functional.cosine_similarity(f, ref, dim=-1)
    assert (1 - cos).max().item() < 1e-5
    assert (f - ref).abs().max().item() < 2e-4


def test_return_all_matches_reference(setup, gold):
    ocfg, sd, img, bm = setup
    with torch.no_grad():
        f = OC.extract_clip(img, bm, sd, ocfg, "fixed_weights", mask_res=336, return_all=True, pool_heads=GG.CROP_POOL_HEADS)
    assert (f - torch.from_numpy(gold["return_all_336"])).abs().max().item() < 2e-4


def test_degenerate_masks_raise_like_the_reference(setup):
    ocfg, sd, img, _ = setup
    bm = np.zeros((1, 480, 640), bool)
    bm[0, 100:200, 77] = True            # one column: w = right - left = 0 -> empty crop -> F.resize raises
    with pytest.raises(RuntimeError):
        OC.extract_clip(img, bm, sd, ocfg, "fixed_weights", mask_res=336, pool_heads=GG.CROP_POOL_HEADS)
    with torch.no_grad():                # `vanilla` pads to a square first: a black image, no error
        f = OC.extract_clip(img, bm, sd, ocfg, "vanilla", mask_res=336, pool_heads=GG.CROP_POOL_HEADS)
    assert f.shape == (1, ocfg.output_dim) and torch.isfinite(f).all()


def test_siglip_similarity():
    g = torch.Generator().manual_seed(0)
    img, txt = torch.randn(5, 16, generator=g), torch.randn(3, 16, generator=g)
    s = OC.siglip_similarity(txt, img, np.log(10.0), -10.0)
    assert torch.allclose(s, torch.sigmoid(img @ txt.T * 10.0 - 10.0), atol=1e-6)


@pytest.mark.parametrize("name", list(GG.MERGER_CFGS))
def test_learned_merger_matches_reference(setup, golden_dir, name):
    """embed_type `learned` (clips_merging.py:26-56): the oracle's merger against the reference module, stand-alone and behind the
    reference's CLIPGenerator."""
    g = np.load(os.path.join(golden_dir, "merger.npz"))
    ocfg, sd, img, bm = setup
    mc = GG.MERGER_CFGS[name]
    msd = GG.merger_state_dict(mc, seed=5)
    gen = torch.Generator().manual_seed(9)
    clips = torch.nn.functional.normalize(torch.randn(11, 3, 64, generator=gen), dim=-1)
    with torch.no_grad():
        m = OC.weights_predictor_merge(clips, msd, nhead=mc["transformer"]["nhead"])
        assert (m - torch.from_numpy(g[f"merge_{name}"])).abs().max().item() < 2e-6
        allc = OC.extract_clip(img, bm, sd, ocfg, "fixed_weights", mask_res=336, return_all=True, pool_heads=GG.CROP_POOL_HEADS)
        f = OC.weights_predictor_merge(allc, msd, nhead=mc["transformer"]["nhead"])
    ref = torch.from_numpy(g[f"learned_{name}"])
    assert (1 - torch.nn.functional.cosine_similarity(f, ref, dim=-1)).max().item() < 1e-5
