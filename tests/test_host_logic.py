"""Host-side logic that needs no GPU: Instance3D view selection, tokenizer, synthetic scenes, sharding."""
import numpy as np
import pytest
import torch

from ovo_b200.instance3d import Instance3D
from ovo_b200 import synth


def test_instance3d_topk_and_to_update():
    Instance3D.n_top_kf = 2
    o = Instance3D(3, kf_id=0, points_ids=[1, 2], mask_area=50)
    assert o.to_update and o.top_kf == [(50, 0)] and o.kfs_ids == [0] and o.points_ids == [1, 2]
    o.to_update = False
    o.update([], 1, 80)
    assert o.to_update and sorted(o.top_kf) == [(50, 0), (80, 1)]
    o.to_update = False
    o.update([], 2, 10)                       # smaller than both: pushed and popped straight away
    assert not o.to_update and not o.is_top_kf(2) and o.kfs_ids == [0, 1, 2]
    o.update([], 3, 60)                       # evicts keyframe 0
    assert o.to_update and not o.is_top_kf(0) and o.is_top_kf(3)
    o.to_update = False
    o.add_top_kf(3, 55)                       # known keyframe, smaller area: nothing changes
    assert not o.to_update
    o.add_top_kf(3, 99)
    assert o.to_update and (99, 3) in o.top_kf
    descs = {1: {3: "d1"}, 3: {3: "d3"}}
    assert o.views_to_fuse(descs) == ["d3", "d1"]          # area-descending (heapq.nlargest)
    assert not o.to_update and o.views_to_fuse(descs) is None
    assert o.views_to_fuse(descs, force_update=True) == ["d3", "d1"]
    Instance3D.n_top_kf = 0
    p = Instance3D(4, kf_id=0, points_ids=[], mask_area=5)
    assert p.to_update and p.top_kf == []                  # n_top_kf <= 0: heap stays empty, flag raised
    assert p.views_to_fuse({0: {4: "x"}}) == ["x"]
    Instance3D.n_top_kf = 10000


def test_instance3d_export_restore_keys():
    o = Instance3D(7, kf_id=2, points_ids=[5], mask_area=9)
    o.clip_feature = torch.ones(4)
    d = o.export(True)
    assert set(d) == {"ins3d_7_clip_feature", "ins3d_7_clip_feature_kf", "ins3d_7_keyframes_ids", "ins3d_7_points_ids", "ins3d_7_top_kfs"}
    q = Instance3D(7)
    q.restore(d, True)
    assert q.kfs_ids == [2] and q.points_ids == [5] and not q.to_update


def test_tokenizer_matches_known_answer():
    from ovo_b200.tokenizer import BPETokenizer, find_vocab
    if find_vocab() is None:
        pytest.skip("CLIP BPE vocabulary not reachable (ships with the reference)")
    t = BPETokenizer(context_length=32)
    ids = t(["a chair"])[0]
    assert ids[:4].tolist() == [49406, 320, 4269, 49407] and int(ids[4:].sum()) == 0
    long = t(["word " * 100])[0]
    assert long[-1].item() == 49407 and long[0].item() == 49406


def test_synth_scene_is_well_formed():
    K = synth.intrinsics()
    d = synth.depth_map()
    assert d.min() == 0 and d[d > 0].min() < d.max()            # non-degenerate frustum
    seg, bm = synth.grid_masks()
    assert seg.max() + 1 == bm.shape[0] == 48 and (seg == -1).any()
    xyz, ids, ins = synth.point_map(1000, d, K, synth.pose(0))
    assert xyz.shape == (1000, 3) and xyz.dtype == np.float32 and (ins == -1).all()


def test_clip_generator_config_validation_needs_no_gpu():
    """Unknown embed types / model cards are rejected before anything touches the device (clip_generator.py:16,37-52)."""
    from ovo_b200.clip_generator import CLIPGenerator, CROP_EMBED_TYPES, MODEL_CARDS
    assert set(CROP_EMBED_TYPES) == {"vanilla", "fixed_weights", "hovsg", "adaptive_weights", "concept_fusion", "learned"}
    assert "PE-Core-L14-336" in MODEL_CARDS and "PE-Core-L-14-336" in MODEL_CARDS      # the vendored name and its open_clip alias
    with pytest.raises(NotImplementedError):
        CLIPGenerator({"embed_type": "no_such_type"})
    with pytest.raises(NotImplementedError):
        CLIPGenerator({"embed_type": "TextRegion", "model_card": "SigLIP-384"})         # open_clip-only card: not built
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                                                # no CPU fallback
            CLIPGenerator({"embed_type": "fixed_weights", "model_card": "PE-Core-L14-336", "random_init": True, "random_init_seed": 0})
    for bad in ({"remove_global_patch": True}, {"resize_method": "resize"}, {"project_and_normalize": False}, {"use_half": True}):
        with pytest.raises(NotImplementedError):      # reference options that are not built are refused (before the device is touched), not ignored
            CLIPGenerator({"embed_type": "TextRegion", "random_init": True, **bad})
    CLIPGenerator._check_supported_keys({"remove_global_patch": False, "resize_method": "multi_resolution", "project_and_normalize": True})


def test_clip_checkpoint_loading_rules(tmp_path):
    """`clip.ckpt_path`: the unwrapping of pe.CLIP.load_ckpt / pe.VisionTransformer.load_ckpt (pe.py:629-638, 407-419) — `state_dict`
    / `weights` wrappers, DDP's `module.` prefix, a vision-only checkpoint without the `visual.` prefix; the reference's root
    directory layout (clip_utils.py:90-93); a missing file or a missing key RAISES; random weights only with `random_init`."""
    from ovo_b200.clip_generator import load_clip_state_dict, normalize_pe_state_dict
    from ovo_b200.encoder import EncoderConfig, random_state_dict
    cfg = EncoderConfig(width=64, layers=2, heads=1, mlp_width=128, output_dim=32, text_width=64, text_heads=1, text_layers=1,
                        text_mlp_width=128, vocab_size=100, text_output_dim=32, image_size=28)
    sd = random_state_dict(cfg, seed=3)
    card = "PE-Core-L14-336"
    # 1. plain state_dict, 2. {"state_dict": ...} with module. prefixes, 3. {"weights": ...}
    torch.save(sd, tmp_path / "a.pt")
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, tmp_path / "b.pt")
    torch.save({"weights": sd}, tmp_path / "c.pt")
    for name in ("a.pt", "b.pt", "c.pt"):
        got = load_clip_state_dict({"ckpt_path": str(tmp_path / name)}, cfg, card)
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    # 4. vision-only checkpoint (VisionTransformer's own keys): gets its `visual.` prefix back, no text tower
    vis = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    torch.save(vis, tmp_path / "v.pt")
    got = load_clip_state_dict({"ckpt_path": str(tmp_path / "v.pt")}, cfg, card)
    assert set(got) == {k for k in sd if k.startswith("visual.")} and "token_embedding.weight" not in got
    assert normalize_pe_state_dict({"state_dict": {"module.visual.proj": 1}}) == {"visual.proj": 1}
    # 5. the reference's directory layout: <root>/data/input/ckpts/pe/<card>.pt
    d = tmp_path / "root" / "data" / "input" / "ckpts" / "pe"
    d.mkdir(parents=True)
    torch.save(sd, d / f"{card}.pt")
    assert set(load_clip_state_dict({"ckpt_path": str(tmp_path / "root")}, cfg, card)) == set(sd)
    # 6. loud failures
    with pytest.raises(FileNotFoundError):
        load_clip_state_dict({"ckpt_path": str(tmp_path / "nope.pt")}, cfg, card)
    with pytest.raises(FileNotFoundError):
        load_clip_state_dict({}, cfg, card)                                # the reference would fail to download: so do we
    torch.save({k: v for k, v in sd.items() if k != "visual.proj"}, tmp_path / "broken.pt")
    with pytest.raises(KeyError):
        load_clip_state_dict({"ckpt_path": str(tmp_path / "broken.pt")}, cfg, card)
    rnd = load_clip_state_dict({"random_init": True, "random_init_seed": 3}, cfg, card)
    assert all(torch.equal(rnd[k], sd[k]) for k in sd)



def test_embed_type_codes_match_the_header():
    import os, re
    from ovo_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ovo_b200.h")).read()
    codes = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define OVO_EMBED_([A-Z_]+) (\d+)", hdr)}
    assert codes == _lib.EMBED_TYPES
