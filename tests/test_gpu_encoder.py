"""GPU parity of the encoder path (E1-E5, Q1) against the CPU oracle and the reference's golden vectors.
Tolerances (north_star: <= 1e-3 cosine vs the reference; SURVEY A6 adds relative-L2 <= 1e-2 because random
weights make cosine a weak discriminator)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder as OE, gen_golden as GG
from ovo_b200 import synth

COS_TOL = 1e-3
REL_TOL = 1e-2


def _ocfg(cfg):
    return OE.VitCfg(**{k: getattr(cfg, k) for k in ("image_size", "patch_size", "width", "layers", "heads", "mlp_width",
                                                      "output_dim", "ln_eps", "text_ctx", "text_width", "text_heads",
                                                      "text_layers", "text_mlp_width", "vocab_size")})


def _check(out, ref, what):
    out, ref = out.float().cpu(), torch.as_tensor(ref).float()
    assert not torch.isnan(out).any(), what
    rel = ((out - ref).norm() / ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(out.reshape(-1, out.shape[-1]), ref.reshape(-1, ref.shape[-1]), dim=-1)
    assert rel < REL_TOL, (what, rel)
    assert (1 - cos).max().item() < COS_TOL, (what, (1 - cos).max().item())


@pytest.fixture(scope="module")
def tiny():
    from ovo_b200.encoder import RegionEncoder, random_state_dict
    cfg = GG.tiny_cfg()
    sd = random_state_dict(cfg, seed=0)
    enc = RegionEncoder(cfg, sd, max_images=8, max_h=968, max_w=1296, max_masks=64)
    return enc, cfg, _ocfg(cfg), sd


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder_tiny.npz"))


def test_tokens_match_reference_golden(tiny, gold):
    """E1 + E2 from uint8 pixels: preprocess (AA resize) + ViT tokens vs the reference's forward_features."""
    enc, cfg, ocfg, sd = tiny
    img = torch.from_numpy(synth.rgb(480, 640, seed=3)).cuda()
    tok = enc.forward_features(img[None])
    assert tok.shape == (2, 577, cfg.width)
    _check(tok[:, ::16], gold["tok_a_sub"], "tokens")


@pytest.mark.parametrize("tag,hw", [("a", (480, 640)), ("b", (968, 1296))])
def test_region_features_match_reference_golden(tiny, gold, tag, hw):
    """E1..E5 (2 images at 480x640, 7 images at 968x1296), incl. the empty mask row."""
    enc, cfg, ocfg, sd = tiny
    img = torch.from_numpy(synth.rgb(*hw, seed=3)).cuda()
    bm = torch.from_numpy(GG.masks_for(*hw)).cuda()
    out = enc.encode_regions(img, bm)
    _check(out, gold[f"regions_{tag}"], f"regions_{tag}")
    assert torch.allclose(out.norm(dim=-1), torch.ones(out.shape[0], device="cuda"), atol=1e-4)


def test_text_matches_reference_golden(tiny, gold):
    enc, cfg, ocfg, sd = tiny
    out = enc.encode_text(torch.from_numpy(GG.TEXT_TOKENS))
    _check(out, gold["text"], "text")


def test_layer_by_layer_vs_oracle(tiny):
    enc, cfg, ocfg, sd = tiny
    torch.manual_seed(1)
    px = torch.randn(3, 3, 336, 336) * 0.5
    for L in (0, 1, 2):
        out = enc.forward_features_from_pixels(px.cuda(), n_layers=L, ln_post=(L == cfg.layers))
        with torch.no_grad():
            ref = OE.vit_forward_features(px, sd, ocfg, n_layers=L, norm=(L == cfg.layers))
        _check(out, ref, f"layer {L}")
    # graph replay (3rd call of a shape) gives the same numbers as the eager first call
    a = enc.forward_features_from_pixels(px.cuda())
    b = enc.forward_features_from_pixels(px.cuda())
    c = enc.forward_features_from_pixels(px.cuda())
    assert torch.equal(a, b) and torch.equal(b, c)


def test_attention_rescale_path(tiny):
    """The attention kernel keeps a row's reference maximum fixed after the first key block and only rescales its TMEM
    accumulator when a later block exceeds it by 2^32 (attention.cuh).  Debug bit 16 forces the rescale whenever the
    maximum grows, so the tcgen05.ld/st rescale path is exercised; both modes must agree with the oracle."""
    from ovo_b200 import _lib
    enc, cfg, ocfg, sd = tiny
    torch.manual_seed(2)
    px = torch.randn(2, 3, 336, 336) * 0.5
    with torch.no_grad():
        ref = OE.vit_forward_features(px, sd, ocfg, n_layers=cfg.layers, norm=True)
        ref1 = OE.vit_forward_features(px[:1], sd, ocfg, n_layers=1, norm=False)
    try:
        _lib.lib().ovo_set_gemm_cluster(16 << 24)
        # (1 image, 1 layer) is a shape no other test uses: its first call runs eagerly with the debug bit
        out = enc.forward_features_from_pixels(px.cuda()[:1], n_layers=1, ln_post=False)
        _check(out, ref1, "forced rescale")
    finally:
        _lib.lib().ovo_set_gemm_cluster(0)
    out2 = enc.forward_features_from_pixels(px.cuda())
    _check(out2, ref, "lazy rescale")


def test_batched_frames_equal_single_frames(tiny):
    """Batching keyframes changes nothing: features of frame f in a batch == features of frame f alone."""
    enc, cfg, ocfg, sd = tiny
    imgs = torch.stack([torch.from_numpy(synth.rgb(480, 640, seed=s)) for s in (1, 2, 3)]).cuda()
    seg, bm = synth.grid_masks(rows=2, cols=3)
    bm = torch.from_numpy(bm).cuda()
    batch = enc.encode_regions(imgs, torch.cat([bm, bm[:2], bm[:4]]), masks_per_frame=[6, 2, 4])
    singles = [enc.encode_regions(imgs[0], bm), enc.encode_regions(imgs[1], bm[:2]), enc.encode_regions(imgs[2], bm[:4])]
    assert torch.equal(batch, torch.cat(singles))


@pytest.fixture(scope="module")
def l14():
    from ovo_b200.encoder import RegionEncoder, EncoderConfig, random_state_dict
    cfg = EncoderConfig(text_layers=2)
    sd = random_state_dict(cfg, seed=0)
    return RegionEncoder(cfg, sd, max_images=4, max_masks=64), cfg, _ocfg(cfg), sd


def test_l14_full_depth_vs_oracle(l14):
    """PE-Core-L14-336 shape, all 24 layers, 2 images, + region pooling, vs the f32 oracle."""
    enc, cfg, ocfg, sd = l14
    img = synth.rgb(480, 640, seed=9)
    seg, bm = synth.grid_masks(rows=3, cols=4)
    out = enc.encode_regions(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda())
    with torch.no_grad():
        ref = OE.encode_regions(img, bm, sd, ocfg)
    _check(out, ref, "L14 regions")
    tok = enc.forward_features(torch.from_numpy(img).cuda()[None])
    with torch.no_grad():
        px = OE.preprocess(torch.from_numpy(img.transpose(2, 0, 1).copy()).float() / 255.0, ocfg)
        ref_tok = OE.vit_forward_features(px, sd, ocfg)
    _check(tok, ref_tok, "L14 tokens")


def test_l14_text_vs_oracle(l14):
    enc, cfg, ocfg, sd = l14
    tok = torch.randint(1, 49000, (21, 32)); tok[:, 12:] = 0; tok[:, 0] = 49406
    tok[torch.arange(21), torch.randint(2, 12, (21,))] = 49407
    out = enc.encode_text(tok)
    with torch.no_grad():
        ref = OE.text_forward(tok, sd, ocfg)
    _check(out, ref, "L14 text")


def test_head_dim_80_generic_attention_path(golden_dir):
    """BASELINE config 4's ViT-H/14 shape has head_dim 80 (1280 / 16): the generic mma.sync attention path
    (generic_attention.cuh) with the rotary embedding applied while staging, vs the oracle and the reference golden."""
    from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict
    cfg = EncoderConfig(**GG.HD80)
    sd = random_state_dict(cfg, seed=0, text=False)
    enc = RegionEncoder(cfg, sd, max_images=8, max_h=968, max_w=1296, max_masks=32)
    torch.manual_seed(4)
    px = torch.randn(2, 3, 336, 336) * 0.5
    ocfg = _ocfg(cfg)
    for L in (1, 2):
        out = enc.forward_features_from_pixels(px.cuda(), n_layers=L, ln_post=(L == cfg.layers))
        with torch.no_grad():
            ref = OE.vit_forward_features(px, sd, ocfg, n_layers=L, norm=(L == cfg.layers))
        _check(out, ref, f"hd80 layer {L}")
    gold = np.load(os.path.join(golden_dir, "encoder_hd80.npz"))
    sub = out.cpu()[:, ::9, ::4].numpy()
    assert np.linalg.norm(sub - gold["tokens_sub"]) / np.linalg.norm(gold["tokens_sub"]) < REL_TOL
    # a 7-image frame (960x1280, config 4's resolution) through the region pipeline
    img = synth.rgb(960, 1280, seed=2)
    seg, bm = synth.grid_masks(960, 1280, rows=2, cols=3)
    feats = enc.encode_regions(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda())
    with torch.no_grad():
        ref = OE.encode_regions(img, bm, sd, ocfg)
    _check(feats, ref, "hd80 regions 960x1280")


def test_vit_h14_shaped_full_depth_vs_oracle():
    """The ViT-H/14-shaped encoder of BASELINE config 4 (width 1280, 32 layers, 16 heads, mlp 5120; 652 M parameters)."""
    from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict
    cfg = EncoderConfig(width=1280, layers=32, heads=16, mlp_width=5120, output_dim=1024, text_layers=0)
    sd = random_state_dict(cfg, seed=3, text=False)
    enc = RegionEncoder(cfg, sd, max_images=2, max_masks=8)
    torch.manual_seed(5)
    px = torch.randn(1, 3, 336, 336) * 0.5
    out = enc.forward_features_from_pixels(px.cuda())
    with torch.no_grad():
        ref = OE.vit_forward_features(px, sd, _ocfg(cfg), n_layers=cfg.layers, norm=True)
    _check(out, ref, "ViT-H/14-shaped, 32 layers")
