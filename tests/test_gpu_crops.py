"""Crop-based descriptors (clip_generator.py:136-158) on the GPU, through the C ABI, against
oracle/crops.py and the reference's own outputs (tests/golden/crops.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import crops as OC, encoder as OE, gen_golden as GG
from ovo_b200 import _lib, synth
from ovo_b200._lib import check, ptr, stream_ptr
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict


def _ocfg(cfg):
    return OE.VitCfg(**{k: getattr(cfg, k) for k in ("image_size", "patch_size", "width", "layers", "heads", "mlp_width",
                                                      "output_dim", "ln_eps", "text_ctx", "text_width", "text_heads",
                                                      "text_layers", "text_mlp_width", "vocab_size")})


@pytest.fixture(scope="module")
def tiny():
    cfg = GG.tiny_cfg()
    sd = random_state_dict(cfg, seed=0)
    enc = RegionEncoder(cfg, sd, max_images=8, max_h=480, max_w=640, max_masks=64, device="cuda:0")
    enc.install_pool_head(sd, pool_heads=GG.CROP_POOL_HEADS)
    return cfg, _ocfg(cfg), sd, enc


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "crops.npz"))


def _boxes(bm):
    m = torch.from_numpy(bm).cuda().to(torch.uint8).contiguous()
    out = torch.empty(m.shape[0], 4, dtype=torch.int32, device="cuda")
    check(_lib.lib().ovo_mask_boxes(ptr(m), m.shape[0], m.shape[1], m.shape[2], ptr(out), stream_ptr()), "ovo_mask_boxes")
    return out.cpu().numpy()


def test_mask_boxes_bit_exact(gold):
    bm = GG.crop_masks()
    assert (_boxes(bm) == gold["boxes_xywh"]).all()
    rng = np.random.default_rng(0)
    odd = np.zeros((5, 37, 53), bool)                 # width not a multiple of 4: the byte-wise path
    for i in range(4):
        y, x = rng.integers(0, 30), rng.integers(0, 45)
        odd[i, y:y + rng.integers(1, 7), x:x + rng.integers(1, 8)] = rng.random((1, 1)) < 2
    odd[3, 36, 52] = True                             # last pixel; mask 4 stays empty -> 0,0,0,0
    assert (_boxes(odd) == OC.mask_boxes_xywh(odd)).all()
    wide = rng.random((3, 480, 640)) < 0.0005         # sparse pixels: every byte lane of the 32-bit path
    assert (_boxes(wide) == OC.mask_boxes_xywh(wide)).all()


@pytest.mark.parametrize("et,res", [("fixed_weights", 336), ("vanilla", 224), ("fixed_weights", 384)])
def test_crops_match_oracle(tiny, et, res):
    cfg, ocfg, sd, enc = tiny
    img, bm = synth.rgb(480, 640, seed=21), GG.crop_masks()
    _, crops = enc.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), et, mask_res=res, return_crops=True)
    ref = OC.seg_images(bm, torch.from_numpy(img.transpose(2, 0, 1).copy()), et != "vanilla", 50, res)   # [M, 3|6, L, L]
    M = bm.shape[0]
    got = crops.cpu().permute(0, 3, 1, 2)                                                                 # [n, 3, L, L]
    ref = ref[:, :3] if et == "vanilla" else torch.cat([ref[:, :3], ref[:, 3:]])
    d = (got.int() - ref.int()).abs()
    # uint8 after a float resize: a value that lands on .5 may round to the other side (weights differ in the last bit)
    assert d.max().item() <= 1 and (d > 0).float().mean().item() < 2e-3
    assert got.shape[0] == (M if et == "vanilla" else 2 * M)


def test_encode_image_head_matches_oracle(tiny, gold):
    cfg, ocfg, sd, enc = tiny
    img = synth.rgb(480, 640, seed=21)
    imt = torch.from_numpy(img.transpose(2, 0, 1).copy()).float() / 255.0
    px = torch.stack([(OE.aa_resize(imt, 336, 336) - 0.5) / 0.5, (OE.aa_resize(imt[:, :300, 100:500], 336, 336) - 0.5) / 0.5])
    with torch.no_grad():
        ref = OC.attn_pool(OE.vit_forward_features(px, sd, ocfg), sd, GG.CROP_POOL_HEADS) @ sd["visual.proj"]
    got = enc.encode_images_from_pixels(px.cuda()).cpu()
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2
    assert (1 - torch.nn.functional.cosine_similarity(got, ref, dim=-1)).max().item() < 1e-3
    g = torch.from_numpy(gold["encode_image_global"])                   # the reference's own encode_image of the frame
    assert (1 - torch.nn.functional.cosine_similarity(got[:1], g, dim=-1)).item() < 1e-3


@pytest.mark.parametrize("et,res", GG.CROP_CASES)
def test_extract_clip_matches_reference_and_oracle(tiny, gold, et, res):
    cfg, ocfg, sd, enc = tiny
    img, bm = synth.rgb(480, 640, seed=21), GG.crop_masks()
    got = enc.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), et, mask_res=res).cpu()
    ref = torch.from_numpy(gold[f"{et}_{res}"])                          # the reference's CLIPGenerator.extract_clip
    assert got.shape == ref.shape
    assert (1 - torch.nn.functional.cosine_similarity(got, ref, dim=-1)).max().item() < 1e-3    # north-star tolerance
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2
    assert (got.norm(dim=-1) - 1).abs().max().item() < 1e-5
    with torch.no_grad():
        orc = OC.extract_clip(img, bm, sd, ocfg, et, mask_res=res, pool_heads=GG.CROP_POOL_HEADS)
    assert (1 - torch.nn.functional.cosine_similarity(got, orc, dim=-1)).max().item() < 1e-3


def test_return_all_and_chunking(tiny, gold):
    """19 images through an encoder that takes 8 (tiny fixture) or 5 at a time: same descriptors."""
    cfg, ocfg, sd, enc = tiny
    img, bm = synth.rgb(480, 640, seed=21), GG.crop_masks()
    allc = enc.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), "fixed_weights", mask_res=336, return_all=True).cpu()
    ref = torch.from_numpy(gold["return_all_336"])
    assert allc.shape == ref.shape == (bm.shape[0], 3, cfg.output_dim)
    assert (1 - torch.nn.functional.cosine_similarity(allc, ref, dim=-1)).max().item() < 1e-3
    enc5 = RegionEncoder(cfg, sd, max_images=5, max_h=480, max_w=640, max_masks=64, device="cuda:0")
    enc5.install_pool_head(sd, pool_heads=GG.CROP_POOL_HEADS)
    a = enc.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), "hovsg", mask_res=336)
    b = enc5.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), "hovsg", mask_res=336)
    assert (a - b).abs().max().item() < 1e-5
    # the TextRegion path still works on the same handle afterwards (the crop path re-uses its job list and tables)
    r1 = enc.encode_regions(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda()).cpu()
    with torch.no_grad():
        r0 = OE.encode_regions(img, bm, sd, ocfg)
    assert (1 - torch.nn.functional.cosine_similarity(r1, r0, dim=-1)).max().item() < 1e-3


@pytest.mark.parametrize("et", [1, 2, 3, 4])
def test_fuse_clips_kernel(et):
    g = torch.Generator().manual_seed(et)
    M, D = 37, 200
    nrm = lambda t: torch.nn.functional.normalize(t, dim=-1)
    cg, seg, bb = nrm(torch.randn(1, D, generator=g)), nrm(torch.randn(M, D, generator=g)), nrm(torch.randn(M, D, generator=g))
    bb = nrm(bb + 0.7 * seg + 0.5 * cg)                       # correlated, like real crops
    name = {v: k for k, v in _lib.EMBED_TYPES.items()}[et]
    ref = OC.fuse_clips(cg.repeat(M, 1), seg, bb, name, 0.4418, 0.1)
    out = torch.empty(M, D, device="cuda")
    a, b, c = cg.cuda().contiguous(), seg.cuda().contiguous(), bb.cuda().contiguous()
    check(_lib.lib().ovo_fuse_clips(ptr(a), ptr(b), ptr(c), M, D, et, 0.4418, 0.1, ptr(out), stream_ptr()), "ovo_fuse_clips")
    assert (out.cpu() - ref).abs().max().item() < 2e-6


def test_siglip_similarity_kernel():
    g = torch.Generator().manual_seed(0)
    img, txt = torch.randn(50, 64, generator=g) / 8, torch.randn(7, 64, generator=g) / 8
    sim = (img @ txt.T).cuda().contiguous()
    check(_lib.lib().ovo_siglip_similarity(ptr(sim), sim.numel(), float(np.log(10.0)), -10.0, stream_ptr()), "siglip")
    assert (sim.cpu() - OC.siglip_similarity(txt, img, np.log(10.0), -10.0)).abs().max().item() < 1e-6


def test_degenerate_and_error_paths(tiny):
    cfg, ocfg, sd, enc = tiny
    img = torch.from_numpy(synth.rgb(480, 640, seed=21)).cuda()
    bm = np.zeros((2, 480, 640), bool)
    bm[0, 50:200, 60:300] = True
    bm[1, 100:200, 77] = True                        # one column: w = right - left = 0, the reference's F.resize raises
    with pytest.raises(RuntimeError, match="empty crop"):
        enc.encode_crops(img, torch.from_numpy(bm).cuda(), "fixed_weights", mask_res=336)
    f = enc.encode_crops(img, torch.from_numpy(bm).cuda(), "vanilla", mask_res=336).cpu()   # padded to a square: black image
    with torch.no_grad():
        ref = OC.extract_clip(synth.rgb(480, 640, seed=21), bm, sd, ocfg, "vanilla", mask_res=336, pool_heads=GG.CROP_POOL_HEADS)
    assert (1 - torch.nn.functional.cosine_similarity(f, ref, dim=-1)).max().item() < 1e-3
    with pytest.raises(RuntimeError):
        enc.encode_crops(img, torch.from_numpy(bm[:1]).cuda(), "fixed_weights", mask_res=4096)
    bare = RegionEncoder(cfg, sd, max_images=2, max_h=480, max_w=640, max_masks=8, device="cuda:0")
    with pytest.raises(RuntimeError, match="install_pool_head"):
        bare.encode_crops(img, torch.from_numpy(bm[:1]).cuda(), "vanilla")
    from ovo_b200 import CLIPGenerator
    with pytest.raises(NotImplementedError):
        CLIPGenerator({"embed_type": "no_such_type"}, state_dict=sd, encoder_config=cfg)
    gen = CLIPGenerator({"embed_type": "hovsg", "mask_res": 336, "max_images": 8, "max_h": 480, "max_w": 640, "max_masks": 16},
                        state_dict=sd, encoder_config=EncoderConfig(**{**GG.TINY, "pool_heads": GG.CROP_POOL_HEADS}))
    assert gen.extract_clip(img.permute(2, 0, 1), torch.zeros(0, 480, 640, dtype=torch.bool, device="cuda")).numel() == 0
    a = gen.extract_clip(img.permute(2, 0, 1), torch.from_numpy(bm[:1]).cuda())             # [3,H,W] like OVO._extract_clip passes
    b = enc.encode_crops(img, torch.from_numpy(bm[:1]).cuda(), "hovsg", mask_res=336)
    assert (a - b).abs().max().item() < 1e-5


def test_full_size_l14_head_dim_128(golden_dir):
    """PE-Core-L14-336 geometry (width 1024, pooler heads 8 -> head_dim 128), 4 layers to keep the CPU oracle short."""
    cfg = EncoderConfig(layers=4, text_layers=0)
    sd = random_state_dict(cfg, seed=3, text=False)
    enc = RegionEncoder(cfg, sd, max_images=8, max_h=480, max_w=640, max_masks=8, device="cuda:0")
    enc.install_pool_head(sd, pool_heads=8)
    img, bm = synth.rgb(480, 640, seed=5), GG.crop_masks()[6:9]
    got = enc.encode_crops(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda(), "adaptive_weights", mask_res=384).cpu()
    ocfg = _ocfg(cfg)
    with torch.no_grad():
        ref = OC.extract_clip(img, bm, sd, ocfg, "adaptive_weights", mask_res=384, pool_heads=8)
    assert (1 - torch.nn.functional.cosine_similarity(got, ref, dim=-1)).max().item() < 1e-3
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2


def test_ovo_api_with_crop_descriptors(tmp_path):
    """The OVO loop with `embed_type: fixed_weights` against the oracle's loop with the oracle's crop descriptors."""
    from oracle.pipeline import OracleOVO
    from ovo_b200 import OVO, CLIPGenerator
    from test_gpu_ovo import _Logger, _TokTokenizer, _replay
    K, xyz, ids, ins, frames = GG.ovo_inputs()
    frames = frames[:2]
    mdir = tmp_path / "masks" / "scene"
    mdir.mkdir(parents=True)
    for f in frames:
        np.save(mdir / f"{f['frame_id']:04d}_seg_map_default.npy", f["seg"])
        np.save(mdir / f"{f['frame_id']:04d}_bmap_default.npy", f["bm"])
    cfg = EncoderConfig(**{**GG.TINY, "pool_heads": GG.CROP_POOL_HEADS})
    sd = random_state_dict(cfg, seed=0)
    config = GG.ovo_config(str(tmp_path / "masks"))
    config["clip"].update({"embed_type": "fixed_weights", "mask_res": 336, "max_images": 16})
    clip = CLIPGenerator(config["clip"], state_dict=sd, tokenizer=_TokTokenizer(), encoder_config=cfg)
    ovo = OVO(config, _Logger(), scene_name="scene", cam_intrinsics=torch.from_numpy(K), clip_generator=clip)
    _replay(ovo, xyz, ids, ins, frames)
    ocfg = _ocfg(cfg)
    orc = OracleOVO(sd, ocfg, K, track_th=config["track_th"], kf_queue_delay=1,
                    encode_fn=lambda image, masks: OC.extract_clip(image, masks, sd, ocfg, "fixed_weights", mask_res=336,
                                                                   pool_heads=GG.CROP_POOL_HEADS))
    pins = ins.copy()
    for f in frames:
        pins = orc.detect_and_track(f["image"], f["depth"], f["c2w"], f["seg"], f["bm"], xyz, pins)
        orc.compute_semantic_info()
    orc.compute_semantic_info(flush=True)
    assert list(ovo.objects.keys()) == list(orc.objects.keys())
    got, ref = ovo.get_objs_clips().cpu(), orc.bank()
    assert (1 - torch.nn.functional.cosine_similarity(got, ref, dim=-1)).max().item() < 1e-3


@pytest.mark.parametrize("name", list(GG.MERGER_CFGS))
def test_learned_merger_matches_reference(tiny, golden_dir, name):
    """embed_type `learned` (clips_merging.py:26-56): the merger alone on seeded descriptors, then behind CLIPGenerator.extract_clip,
    against the reference module / the reference's CLIPGenerator (tests/golden/merger.npz)."""
    from ovo_b200 import CLIPGenerator
    from ovo_b200.clip_generator import LearnedMerger
    g = np.load(os.path.join(golden_dir, "merger.npz"))
    cfg, ocfg, sd, enc = tiny
    mc = GG.MERGER_CFGS[name]
    msd = GG.merger_state_dict(mc, seed=5)
    gen = torch.Generator().manual_seed(9)
    clips = torch.nn.functional.normalize(torch.randn(11, 3, 64, generator=gen), dim=-1)
    merger = LearnedMerger(mc, msd, "cuda:0")
    got = merger(clips.cuda()).cpu()
    ref = torch.from_numpy(g[f"merge_{name}"])
    assert (1 - torch.nn.functional.cosine_similarity(got, ref, dim=-1)).max().item() < 1e-3
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2
    with torch.no_grad():
        orc = OC.weights_predictor_merge(clips, msd, nhead=mc["transformer"]["nhead"])
    assert (1 - torch.nn.functional.cosine_similarity(got, orc, dim=-1)).max().item() < 1e-3
    clipgen = CLIPGenerator({"embed_type": "learned", "mask_res": 336}, encoder=enc, merger_config=mc, merger_state_dict=msd)
    img, bm = synth.rgb(480, 640, seed=21), GG.crop_masks()
    f = clipgen.extract_clip(torch.from_numpy(img).cuda(), torch.from_numpy(bm).cuda()).cpu()
    ref = torch.from_numpy(g[f"learned_{name}"])
    assert f.shape == ref.shape
    assert (1 - torch.nn.functional.cosine_similarity(f, ref, dim=-1)).max().item() < 1e-3
    assert ((f - ref).norm() / ref.norm()).item() < 1e-2
    with pytest.raises(ValueError):          # a predictor trained for another descriptor width (the reference's is SigLIP's 1152)
        CLIPGenerator({"embed_type": "learned"}, encoder=enc, merger_state_dict=msd,
                      merger_config={"transformer": {**mc["transformer"], "d_model": 1152}, "mlp": mc["mlp"]})


def test_learned_merger_reference_geometry():
    """The shipped hparams (data/input/weights_predictor/base/hparams.yaml: d_model 1152, 8 heads, 5 layers, MLP 3456 -> 13824 x5 ->
    3456) with seeded weights, against the oracle."""
    from ovo_b200.clip_generator import LearnedMerger
    mc = {"transformer": {"d_model": 1152, "nhead": 8, "dim_feedforward": 1152, "n_layers": 5},
          "mlp": {"i_dim": 3456, "h_dim": 13824, "o_dim": 3456, "n_layers": 4, "act_key": "leaky_relu"}}
    msd = GG.merger_state_dict(mc, seed=2)
    gen = torch.Generator().manual_seed(4)
    clips = torch.nn.functional.normalize(torch.randn(40, 3, 1152, generator=gen) + torch.randn(40, 1, 1152, generator=gen), dim=-1)
    got = LearnedMerger(mc, msd, "cuda:0")(clips.cuda()).cpu()
    with torch.no_grad():
        orc = OC.weights_predictor_merge(clips, msd, nhead=8)
    assert (1 - torch.nn.functional.cosine_similarity(got, orc, dim=-1)).max().item() < 1e-3
    assert (got.norm(dim=-1) - 1).abs().max().item() < 1e-5
