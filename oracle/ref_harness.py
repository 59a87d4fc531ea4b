"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (tberriel/OVO @ /root/reference) on CPU.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
`oracle/gen_golden.py` to produce the fixtures under `tests/golden/` and by the `not gpu` tests
that pin the restatements in `oracle/*.py` against the reference when it is present.

Nothing in the product path (`ovo_b200/`) imports this file.

Shims (oracle/shims/) stand in for packages the reference imports but that are absent here and are
not on the executed path: timm.layers.DropPath (pe.py:14, identity at drop_path=0), ftfy.fix_text
(tokenizer.py:14, identity on ASCII), open_clip / open3d / imageio (imported, never called on the
default TextRegion path), hydra / omegaconf / iopath (SAM-2 package import only).
"""
import os
import sys

REF_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ovo"))


def setup_paths() -> None:
    paths = [
        os.path.join(_HERE, "shims"),
        os.path.join(REF_ROOT, "thirdParty/perception_models"),
        os.path.join(REF_ROOT, "thirdParty/segment-anything-2"),
        REF_ROOT,
    ]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)


def tiny_vision_cfg(width=128, layers=2, heads=2, image_size=336, output_dim=64):
    """A small PE config (same code path as PE-Core-L14-336: cls token, abs pos-emb, 2D RoPE,
    attention pooler) so golden fixtures stay small."""
    setup_paths()
    from core.vision_encoder.config import PEConfig
    return PEConfig(image_size=image_size, patch_size=14, width=width, layers=layers, heads=heads,
                    mlp_ratio=4.0, pool_type="attn", output_dim=output_dim, use_cls_token=True,
                    attn_pooler_heads=heads)


def tiny_text_cfg(width=128, layers=2, heads=2, output_dim=64, context_length=32):
    setup_paths()
    from core.vision_encoder.config import PETextConfig
    return PETextConfig(context_length=context_length, width=width, heads=heads, layers=layers,
                        output_dim=output_dim)


def build_clip(vision_cfg=None, text_cfg=None, seed=0, card="PE-Core-L14-336"):
    """pe.CLIP with seeded random weights (no checkpoints on disk, no network).  The two text
    parameters the reference leaves as torch.empty (pe.py:585,619) are initialised explicitly.
    Biases / LN affine are randomised too so that a dropped bias shows up in parity tests."""
    setup_paths()
    import torch
    import core.vision_encoder.pe as pe
    from core.vision_encoder.config import PE_VISION_CONFIG, PE_TEXT_CONFIG
    torch.manual_seed(seed)
    vcfg = vision_cfg or PE_VISION_CONFIG[card]
    tcfg = text_cfg or PE_TEXT_CONFIG[card]
    model = pe.CLIP(vcfg, tcfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        model.positional_embedding.copy_(0.01 * torch.randn(model.positional_embedding.shape, generator=g))
        model.text_projection.copy_(tcfg.width ** -0.5 * torch.randn(model.text_projection.shape, generator=g))
        model.token_embedding.weight.copy_(0.02 * torch.randn(model.token_embedding.weight.shape, generator=g))
        for name, p in model.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif (".ln_" in name or "layernorm" in name or "ln_final" in name or "ln_pre" in name
                  or "ln_post" in name) and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
    return model


def build_textregion(model, card="PE-Core-L14-336"):
    """PETextRegion wired the way CLIPGenerator does it (clip_generator.py:40-49) but on CPU."""
    setup_paths()
    from torchvision.transforms import Resize, Normalize, CenterCrop, Compose
    import core.vision_encoder.transforms as transforms
    from ovo.entities.textregion import PETextRegion
    pre = transforms.get_image_transform(model.image_size)
    keep = [tf for tf in pre.transforms if isinstance(tf, (Resize, CenterCrop, Normalize))]
    tr = PETextRegion(model, model_card=card, preprocess=Compose(keep), resize_method="multi_resolution",
                      remove_global_patch=False, project_and_normalize=True, device="cpu")
    return tr


def tokenizer(context_length=32):
    setup_paths()
    import core.vision_encoder.transforms as transforms
    return transforms.get_text_tokenizer(context_length)


def build_sam2_reference(cfg, state_dict):
    """The reference's SAM2Base (image path) instantiated from ITS OWN yaml (sam2.1_hiera_l.yaml) with the trunk
    geometry of `cfg` (ovo_b200.sam_config.SamConfig) and the given state_dict loaded.  hydra is absent here, so the
    `_target_` tree is instantiated by the 12-line importer below.  The memory modules the
    image path never calls keep their default initialisation (strict=False)."""
    setup_paths()
    import importlib
    import yaml

    path = os.path.join(REF_ROOT, "thirdParty/segment-anything-2/sam2/configs/sam2.1/sam2.1_hiera_l.yaml")
    with open(path) as f:
        y = yaml.safe_load(f)["model"]
    tr = y["image_encoder"]["trunk"]
    tr["embed_dim"], tr["num_heads"] = cfg.embed_dim, cfg.num_heads
    tr["stages"], tr["global_att_blocks"] = list(cfg.stages), list(cfg.global_att_blocks)
    tr["window_spec"] = list(cfg.window_spec)
    y["image_encoder"]["neck"]["backbone_channel_list"] = cfg.channel_list()
    y["image_size"] = cfg.image_size

    def inst(node):
        if isinstance(node, dict):
            kw = {k: inst(v) for k, v in node.items() if k != "_target_"}
            if "_target_" in node:
                mod, name = node["_target_"].rsplit(".", 1)
                return getattr(importlib.import_module(mod), name)(**kw)
            return kw
        if isinstance(node, list):
            return [inst(v) for v in node]
        if isinstance(node, str):
            try:
                return float(node) if any(c in node for c in "eE.") and node.replace(".", "").replace("e", "").replace("E", "").replace("-", "").replace("+", "").isdigit() else node
            except ValueError:
                return node
        return node

    model = inst(y).eval()
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    used = ("image_encoder.", "sam_prompt_encoder.pe_layer", "sam_prompt_encoder.point_embeddings",
            "sam_prompt_encoder.not_a_point", "sam_prompt_encoder.no_mask", "sam_mask_decoder.", "no_mem_embed")
    bad = [k for k in missing if k.startswith(used)]
    assert not bad, f"image-path parameters missing from the state_dict: {bad[:5]}"
    return model


def build_sam2_amg(model, **kw):
    """SAM2AutomaticMaskGenerator wired as ovo/utils/segment_utils.py:296-307 does (points_per_side 16 from ovo.yaml:32)."""
    setup_paths()
    from sam2.automatic_mask_generator import SAM2AutomaticMaskGenerator
    args = dict(points_per_side=16, pred_iou_thresh=0.8, stability_score_thresh=0.95, min_mask_region_area=0, use_m2m=False)
    args.update(kw)
    return SAM2AutomaticMaskGenerator(model=model, **args)
