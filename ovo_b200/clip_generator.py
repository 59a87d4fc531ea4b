"""`CLIPGenerator` with the reference's surface (ovo/entities/clip_generator.py:12-198) on the sm_100a encoder:
the TextRegion branch (`embed_type: TextRegion`, the default of data/working/configs/ovo.yaml:45) and the crop-based
branch (`vanilla`, `fixed_weights`, `hovsg`, `adaptive_weights`, `concept_fusion`; clip_generator.py:136-158) on the
Perception-Encoder card (`PE-Core-L-14-336` in the reference's open_clip table, clip_utils.py:61 — the same
architecture as the vendored `PE-Core-L14-336`).  `embed_type: learned` (a SigLIP-1152 descriptor merger with its own
checkpoint, clips_merging.py) and the open_clip-only cards are not built.

Weights: `config["ckpt_path"]` points to a Perception-Encoder checkpoint (`torch.save`d state_dict with the
reference's key names, what `pe.CLIP.load_ckpt` reads).  A missing checkpoint raises; `config["random_init"]: True`
asks for seeded random weights of the same architecture (benchmarks and tests: there is no network here)."""
import os
from typing import Dict, List

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .encoder import EncoderConfig, RegionEncoder, random_state_dict


class LearnedMerger:
    """`WeightsPredictorMerger` (ovo/entities/clips_merging.py:26-56) on the device: packs the module's state_dict
    (`att_encoder.layers.{l}.*`, `mlp.{2j}.*`) and calls `ovo_merge_clips_learned`."""

    def __init__(self, model_config: Dict, state_dict: Dict, device):
        import ctypes as C
        from ._lib import MergerLayer, MergerWeights
        self.device = torch.device(device)
        self._keep = []
        t = model_config["transformer"]
        f32, bf = torch.float32, torch.bfloat16

        def dev(x, dt):
            x = x.detach().to(self.device, dt).contiguous()
            self._keep.append(x)
            return ptr(x)
        n_layers = t["n_layers"]
        layers = (MergerLayer * max(n_layers, 1))()
        for l in range(n_layers):
            p = f"att_encoder.layers.{l}."
            layers[l] = MergerLayer(dev(state_dict[p + "self_attn.in_proj_weight"], bf), dev(state_dict[p + "self_attn.in_proj_bias"], f32),
                                    dev(state_dict[p + "self_attn.out_proj.weight"], bf), dev(state_dict[p + "self_attn.out_proj.bias"], f32),
                                    dev(state_dict[p + "norm1.weight"], f32), dev(state_dict[p + "norm1.bias"], f32),
                                    dev(state_dict[p + "linear1.weight"], bf), dev(state_dict[p + "linear1.bias"], f32),
                                    dev(state_dict[p + "linear2.weight"], bf), dev(state_dict[p + "linear2.bias"], f32),
                                    dev(state_dict[p + "norm2.weight"], f32), dev(state_dict[p + "norm2.bias"], f32))
        idx = sorted(int(k.split(".")[1]) for k in state_dict if k.startswith("mlp.") and k.endswith(".weight"))
        n = len(idx)
        mw, mb, mo = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int * n)()
        for j, i in enumerate(idx):
            mw[j], mb[j] = dev(state_dict[f"mlp.{i}.weight"], bf), dev(state_dict[f"mlp.{i}.bias"], f32)
            mo[j] = state_dict[f"mlp.{i}.weight"].shape[0]
        self._keep += [layers, mw, mb, mo]
        self.d_model = t["d_model"]
        self.w = MergerWeights(t["d_model"], t.get("nhead", 8), t["dim_feedforward"], n_layers, layers, n, mw, mb, mo, 1e-5)

    def __call__(self, clips: torch.Tensor) -> torch.Tensor:
        """clips [B,3,D] -> [B,D] unit norm."""
        import ctypes as C
        clips = clips.to(self.device, torch.float32).contiguous()
        assert clips.dim() == 3 and clips.shape[1] == 3 and clips.shape[2] == self.d_model, "merger expects [B, 3, d_model] descriptors"
        out = torch.empty(clips.shape[0], clips.shape[2], device=self.device, dtype=torch.float32)
        if clips.shape[0]:
            check(_lib.lib().ovo_merge_clips_learned(C.byref(self.w), ptr(clips), clips.shape[0], ptr(out), stream_ptr()),
                  "ovo_merge_clips_learned")
        return out


def normalize_pe_state_dict(sd: Dict) -> Dict:
    """The unwrapping `pe.CLIP.load_ckpt` / `pe.VisionTransformer.load_ckpt` do (pe.py:629-638, 407-419): a `state_dict` or
    `weights` wrapper, DDP's `module.` prefix.  A vision-only checkpoint (keys of `VisionTransformer` itself, which that loader
    reads after stripping `visual.`) is accepted too: its keys get the `visual.` prefix back, the text tower stays absent."""
    if "state_dict" in sd:
        sd = sd["state_dict"]
    elif "weights" in sd:
        sd = sd["weights"]
    sd = {k.replace("module.", ""): v for k, v in sd.items()}
    if not any(k.startswith("visual.") for k in sd) and "conv1.weight" in sd:
        sd = {"visual." + k: v for k, v in sd.items()}
    return sd


def load_clip_state_dict(config: Dict, cfg: EncoderConfig, model_card: str) -> Dict:
    """Weights of the card: `config["ckpt_path"]` (a file, or the reference's root under which
    `data/input/ckpts/pe/{card}.pt` lives, clip_utils.py:90-93).  A missing checkpoint RAISES, as the reference's loader
    does; seeded random weights are only used on request (`random_init: True`: benchmarks and tests, no network here)."""
    ckpt = config.get("ckpt_path")
    if ckpt:
        ckpt = str(ckpt)
        if os.path.isdir(ckpt):
            ckpt = os.path.join(ckpt, "data", "input", "ckpts", "pe", f"{model_card}.pt")
        if not os.path.exists(ckpt):
            raise FileNotFoundError(f"ovo_b200: checkpoint {ckpt} of {model_card} not found")
        sd = normalize_pe_state_dict(torch.load(ckpt, map_location="cpu", weights_only=True))
        missing = [k for k in ("visual.conv1.weight", "visual.positional_embedding", "visual.proj",
                               "visual.attn_pool.attn.in_proj_weight", f"visual.transformer.resblocks.{cfg.layers - 1}.mlp.c_proj.weight")
                   if k not in sd]
        if missing:
            raise KeyError(f"ovo_b200: checkpoint {ckpt} lacks {missing} (expected the key names of pe.CLIP / pe.VisionTransformer)")
        return sd
    if config.get("random_init", False):
        seed = int(config.get("random_init_seed", 0))
        print(f"[ovo_b200] random_init: seeded RANDOM weights for {model_card} (seed {seed}) — descriptors carry no semantics")
        return random_state_dict(cfg, seed=seed)
    raise FileNotFoundError(f"ovo_b200: no checkpoint for {model_card}: set clip.ckpt_path (the reference downloads it, there is no "
                            "network here) or clip.random_init: True for seeded random weights")


MODEL_CARDS = {"PE-Core-L14-336": EncoderConfig(), "PE-Core-L-14-336": EncoderConfig()}
CROP_EMBED_TYPES = ("vanilla", "fixed_weights", "hovsg", "adaptive_weights", "concept_fusion", "learned")


class CLIPGenerator:
    def __init__(self, config: Dict, device: str = "cuda", state_dict: dict | None = None, tokenizer=None,
                 encoder_config: EncoderConfig | None = None, encoder: RegionEncoder | None = None,
                 merger_config: Dict | None = None, merger_state_dict: Dict | None = None):
        self.config = config
        self.device = device
        self.embed_type = config.get("embed_type", "vanilla")
        self.mask_res = config.get("mask_res", 384)
        if self.embed_type != "TextRegion" and self.embed_type not in CROP_EMBED_TYPES:
            raise NotImplementedError(
                f"ovo_b200: embed_type '{self.embed_type}' is not built (have TextRegion, {', '.join(CROP_EMBED_TYPES)})")
        self._check_supported_keys(config)
        self.w_masked = config.get("w_masked", 0.4418)      # clip_generator.py:33-34
        self.w_global = config.get("w_global", 0.1)
        self.model_card = config.get("model_card", "PE-Core-L14-336")
        cfg = (encoder.cfg if encoder is not None else None) or encoder_config or MODEL_CARDS.get(self.model_card)
        if cfg is None:
            raise NotImplementedError(f"ovo_b200: model card '{self.model_card}' is not supported (have {list(MODEL_CARDS)})")
        if encoder is None and not torch.cuda.is_available():
            raise RuntimeError("ovo_b200.CLIPGenerator needs a CUDA device (there is no CPU fallback)")
        if state_dict is None and encoder is None:
            state_dict = load_clip_state_dict(config, cfg, self.model_card)
        self.cfg = cfg
        self.clip_dim = cfg.output_dim
        self.encoder = encoder or RegionEncoder(cfg, state_dict, max_images=config.get("max_images", 16),
                                                max_h=config.get("max_h", 1080), max_w=config.get("max_w", 1920),
                                                max_masks=config.get("max_masks", 512), device=device)
        self.model = self.encoder                      # attribute the reference exposes
        self._tokenizer = tokenizer
        if self.embed_type in CROP_EMBED_TYPES and not self.encoder.has_pool_head:
            self.encoder.install_pool_head(state_dict, pool_heads=getattr(cfg, "pool_heads", 8))
        self.clips_fusion_model = None
        if self.embed_type == "learned":          # clip_generator.py:18-29: hparams.yaml + model.pt under weights_predictor_path
            import yaml
            mcfg, msd = merger_config, merger_state_dict
            if msd is None:
                with open(os.path.join(config["weights_predictor_path"], "hparams.yaml"), "r") as f:
                    mcfg = yaml.safe_load(f)["model"]
                msd = torch.load(os.path.join(config["weights_predictor_path"], "model.pt"), map_location="cpu", weights_only=True)
            if mcfg["transformer"]["d_model"] != self.clip_dim:
                raise ValueError(f"weights predictor was trained for {mcfg['transformer']['d_model']}-d descriptors (the reference's SigLIP "
                                 f"cards), the encoder emits {self.clip_dim}")
            self.clips_fusion_model = LearnedMerger(mcfg, msd, self.encoder.device)
        # clip_generator.py:54-72: SigLIP cards score with sigmoid(sim * exp(logit_scale) + logit_bias).  No SigLIP
        # architecture is built in; the rule applies when a caller supplies such an encoder_config under a SigLIP card.
        self.similarity_args = ()
        if self.model_card.startswith("SigLIP"):
            if config.get("logit_scale") is None or config.get("logit_bias") is None:
                raise NotImplementedError("SigLIP cards need `logit_scale` and `logit_bias` in the config")
            self.similarity_args = (float(config["logit_scale"]), float(config["logit_bias"]))
        self._text_cache = {}

    @staticmethod
    def _check_supported_keys(config: Dict) -> None:
        """Reference options that change the descriptors and are not built: refuse them instead of ignoring them
        (clip_generator.py:30,45-47; ovo.py:437)."""
        bad = []
        if config.get("remove_global_patch", False):
            bad.append("remove_global_patch: True")
        if config.get("resize_method", "multi_resolution") != "multi_resolution":
            bad.append(f"resize_method: {config['resize_method']}")
        if not config.get("project_and_normalize", True):
            bad.append("project_and_normalize: False")
        if config.get("use_half", False):
            bad.append("use_half: True")
        if bad:
            raise NotImplementedError("ovo_b200: clip option(s) not built: " + ", ".join(bad))

    @property
    def get_clip_dim(self) -> int:
        return self.clip_dim

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from .tokenizer import BPETokenizer
            self._tokenizer = BPETokenizer(context_length=self.cfg.text_ctx)
        return self._tokenizer

    def to(self, device: str) -> None:
        return self.cuda() if "cuda" in device else self.cpu()

    def cpu(self) -> None:
        """The reference parks the model on the host to free VRAM; weights here stay device-resident (0.7 GB)."""
        self.device = "cpu"

    def cuda(self) -> None:
        self.device = "cuda"

    @torch.no_grad()
    def extract_clip(self, image: torch.Tensor, binary_maps: torch.Tensor, return_all: bool = False) -> torch.Tensor:
        """image [3,H,W] (or [H,W,3] uint8) range 0-255, binary_maps [N,H,W] -> [N, clip_dim] on the device
        (clip_generator.py:125-158); [N,3,clip_dim] with return_all on the crop-based types."""
        if image.dim() == 3 and image.shape[0] == 3 and image.shape[-1] != 3:
            image = image.permute(1, 2, 0)
        image = image.to(self.encoder.device).round().to(torch.uint8) if image.dtype != torch.uint8 else image
        if self.embed_type == "TextRegion":
            return self.encoder.encode_regions(image.contiguous(), binary_maps)
        if binary_maps.shape[0] == 0:
            return torch.tensor([], device=self.encoder.device)          # clip_generator.py:141-142
        if self.embed_type == "learned":          # clip_generator.py:150-154: the three descriptors, merged by the predictor
            allc = self.encoder.encode_crops(image.contiguous(), binary_maps, "fixed_weights", mask_res=self.mask_res, return_all=True)
            return allc if return_all else self.clips_fusion_model(allc)
        return self.encoder.encode_crops(image.contiguous(), binary_maps, self.embed_type, mask_res=self.mask_res,
                                         w_masked=self.w_masked, w_global=self.w_global, return_all=return_all)

    def _tokens(self, phrases: List[str]) -> torch.Tensor:
        return torch.cat([self.tokenizer(p) for p in phrases])

    @torch.no_grad()
    def get_txt_embedding(self, text_list: List[str]) -> torch.Tensor:
        """L2-normalised text embeddings, one per string (clip_generator.py:161-174)."""
        return self.text_bank([[t] for t in text_list])

    @torch.no_grad()
    def text_bank(self, queries: List[List[str]]) -> torch.Tensor:
        """queries[q] = the template-expanded phrases of query q -> [Q, D] = normalize(mean_t(normalize(e)))
        (clip_generator.py:191-196), cached per phrase tuple."""
        key = tuple(tuple(q) for q in queries)
        if key not in self._text_cache:
            T = len(queries[0])
            assert all(len(q) == T for q in queries)
            tok = self._tokens([p for q in queries for p in q]).to(self.encoder.device, torch.int32).contiguous()
            out = torch.empty(len(queries), self.cfg.text_output_dim, device=self.encoder.device, dtype=torch.float32)
            check(self.encoder.lib.ovo_text_bank(self.encoder.handle, ptr(tok), len(queries), T, ptr(out), stream_ptr()),
                  "ovo_text_bank")
            self._text_cache[key] = out
        return self._text_cache[key]

    @torch.no_grad()
    def get_embed_txt_similarity(self, ins_descriptors: torch.Tensor, txt_queries: List[str],
                                 templates: str | List[str] = ['{}'], rows: torch.Tensor | None = None) -> torch.Tensor:
        """[n_obj, n_queries] plain dot products — PE cards use no logit scale (clip_generator.py:54-72,176-199)."""
        if isinstance(templates, str):
            templates = [templates]
        queries = [[t.format(q) for t in templates] for q in txt_queries]
        txt = self.text_bank(queries)
        bank = ins_descriptors.to(self.encoder.device, torch.float32).contiguous()
        n = bank.shape[0] if rows is None else rows.shape[0]
        out = torch.empty(n, txt.shape[0], device=self.encoder.device, dtype=torch.float32)
        check(_lib.lib().ovo_query_instances(ptr(bank), ptr(rows), n, bank.shape[1], ptr(txt), txt.shape[0], ptr(out),
                                             stream_ptr()), "ovo_query_instances")
        if self.similarity_args and out.numel() > 0:        # clip_utils.py:10-14
            check(_lib.lib().ovo_siglip_similarity(ptr(out), out.numel(), self.similarity_args[0], self.similarity_args[1],
                                                   stream_ptr()), "ovo_siglip_similarity")
        return out
