"""Host side of the region/text encoder: packs a Perception-Encoder state_dict (the reference's own key
names) into device buffers and drives libovo_b200 through the C ABI.

Mirrors what `PETextRegion` (ovo/entities/textregion.py:52-203) and `pe.CLIP.encode_text`
(thirdParty/perception_models/core/vision_encoder/pe.py:725) do in the reference."""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import VitCfg, VitWeights, BlockWeights, PoolHeadWeights, CropParams, EMBED_TYPES, check, ptr, stream_ptr


@dataclass
class EncoderConfig:
    """PE-Core-L14-336 defaults (config.py:101-118)."""
    image_size: int = 336
    patch_size: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp_width: int = 4096
    output_dim: int = 1024
    ln_eps: float = 1e-5
    text_ctx: int = 32
    text_width: int = 1024
    text_heads: int = 16
    text_layers: int = 24
    text_mlp_width: int = 4096
    vocab_size: int = 49408
    text_output_dim: int = 1024
    pool_heads: int = 8          # attn_pooler_heads of the attention-pooling head (config.py:46)

    @property
    def grid(self):
        return self.image_size // self.patch_size

    @property
    def seq(self):
        return self.grid * self.grid + 1


def random_state_dict(cfg: EncoderConfig, seed: int = 0, text: bool = True, device="cpu"):
    """Seeded random weights with the reference's key names and shapes (no checkpoints are available
    offline).  Scales follow pe.py's init so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    W, F, D = cfg.width, cfg.mlp_width, cfg.output_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}

    def block(pfx, W, F):
        sd[pfx + "ln_1.weight"] = 1 + rn(W, std=0.05); sd[pfx + "ln_1.bias"] = rn(W, std=0.02)
        sd[pfx + "attn.in_proj_weight"] = rn(3 * W, W, std=W ** -0.5); sd[pfx + "attn.in_proj_bias"] = rn(3 * W, std=0.02)
        sd[pfx + "attn.out_proj.weight"] = rn(W, W, std=W ** -0.5); sd[pfx + "attn.out_proj.bias"] = rn(W, std=0.02)
        sd[pfx + "ln_2.weight"] = 1 + rn(W, std=0.05); sd[pfx + "ln_2.bias"] = rn(W, std=0.02)
        sd[pfx + "mlp.c_fc.weight"] = rn(F, W, std=W ** -0.5); sd[pfx + "mlp.c_fc.bias"] = rn(F, std=0.02)
        sd[pfx + "mlp.c_proj.weight"] = rn(W, F, std=F ** -0.5); sd[pfx + "mlp.c_proj.bias"] = rn(W, std=0.02)

    sd["visual.conv1.weight"] = rn(W, 3, cfg.patch_size, cfg.patch_size, std=(3 * cfg.patch_size ** 2) ** -0.5)
    sd["visual.class_embedding"] = rn(W, std=W ** -0.5)
    sd["visual.positional_embedding"] = rn(cfg.seq, W, std=W ** -0.5)
    for n in ("ln_pre", "ln_post"):
        sd[f"visual.{n}.weight"] = 1 + rn(W, std=0.05); sd[f"visual.{n}.bias"] = rn(W, std=0.02)
    for i in range(cfg.layers):
        block(f"visual.transformer.resblocks.{i}.", W, F)
    sd["visual.proj"] = rn(W, D, std=W ** -0.5)
    sd["visual.attn_pool.attn.in_proj_weight"] = rn(3 * W, W, std=W ** -0.5)
    sd["visual.attn_pool.attn.in_proj_bias"] = rn(3 * W, std=0.02)
    sd["visual.attn_pool.attn.out_proj.weight"] = rn(W, W, std=W ** -0.5)
    sd["visual.attn_pool.attn.out_proj.bias"] = rn(W, std=0.02)
    if text and cfg.text_layers > 0:
        TW, TF = cfg.text_width, cfg.text_mlp_width
        sd["token_embedding.weight"] = rn(cfg.vocab_size, TW, std=0.02)
        sd["positional_embedding"] = rn(cfg.text_ctx, TW, std=0.01)
        for i in range(cfg.text_layers):
            block(f"transformer.resblocks.{i}.", TW, TF)
        sd["ln_final.weight"] = 1 + rn(TW, std=0.05); sd["ln_final.bias"] = rn(TW, std=0.02)
        sd["text_projection"] = rn(TW, cfg.text_output_dim, std=TW ** -0.5)
    # attention-pooling head of `encode_image` (pe.py:44-87), used by the crop-based embed types.  Drawn LAST so the
    # values of every key above (and the goldens generated from them) do not depend on it.
    p = "visual.attn_pool."
    sd[p + "probe"] = rn(1, 1, W)
    sd[p + "layernorm.weight"] = 1 + rn(W, std=0.05); sd[p + "layernorm.bias"] = rn(W, std=0.02)
    PF = 4 * W                                          # AttentionPooling's own mlp_ratio = 4 (pe.py:52)
    sd[p + "mlp.c_fc.weight"] = rn(PF, W, std=W ** -0.5); sd[p + "mlp.c_fc.bias"] = rn(PF, std=0.02)
    sd[p + "mlp.c_proj.weight"] = rn(W, PF, std=PF ** -0.5); sd[p + "mlp.c_proj.bias"] = rn(W, std=0.02)
    return {k: v.to(device) for k, v in sd.items()}


class RegionEncoder:
    """Device-resident PE ViT + TextRegion pooling + text tower."""

    def __init__(self, cfg: EncoderConfig, state_dict: dict, max_images: int = 16, max_h: int = 480, max_w: int = 640,
                 max_masks: int = 256, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("ovo_b200.RegionEncoder needs a CUDA device (no CPU fallback)")
        self.cfg, self.device = cfg, torch.device(device)
        self.lib = _lib.lib()
        self._keep = []
        self.has_text = cfg.text_layers > 0 and "token_embedding.weight" in state_dict
        w = self._pack(state_dict)
        c = VitCfg(cfg.image_size, cfg.patch_size, cfg.width, cfg.layers, cfg.heads, cfg.mlp_width, cfg.output_dim,
                   cfg.ln_eps, cfg.text_ctx, cfg.text_width, cfg.text_heads, cfg.text_layers if self.has_text else 0,
                   cfg.text_mlp_width, cfg.vocab_size, cfg.text_output_dim)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ovo_encoder_create(C.byref(c), C.byref(w), max_images, max_h, max_w, max_masks, C.byref(h)),
                  "ovo_encoder_create")
        self.handle = h
        self.max_images, self.max_masks = max_images, max_masks
        self.has_pool_head = False
        self._sd_for_head = ({k: v for k, v in state_dict.items() if k.startswith("visual.attn_pool.") or k == "visual.proj"}
                             if "visual.attn_pool.probe" in state_dict else None)

    # ------------------------------------------------------------------ weights
    def _dev(self, t, dtype):
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(t)
        return t

    def _blocks(self, sd, pfx, n):
        arr = (BlockWeights * max(n, 1))()
        for i in range(n):
            p = f"{pfx}{i}."
            f32, bf = torch.float32, torch.bfloat16
            arr[i] = BlockWeights(
                ptr(self._dev(sd[p + "ln_1.weight"], f32)), ptr(self._dev(sd[p + "ln_1.bias"], f32)),
                ptr(self._dev(sd[p + "attn.in_proj_weight"], bf)), ptr(self._dev(sd[p + "attn.in_proj_bias"], f32)),
                ptr(self._dev(sd[p + "attn.out_proj.weight"], bf)), ptr(self._dev(sd[p + "attn.out_proj.bias"], f32)),
                ptr(self._dev(sd[p + "ln_2.weight"], f32)), ptr(self._dev(sd[p + "ln_2.bias"], f32)),
                ptr(self._dev(sd[p + "mlp.c_fc.weight"], bf)), ptr(self._dev(sd[p + "mlp.c_fc.bias"], f32)),
                ptr(self._dev(sd[p + "mlp.c_proj.weight"], bf)), ptr(self._dev(sd[p + "mlp.c_proj.bias"], f32)))
        self._keep.append(arr)
        return arr

    def _pack(self, sd) -> VitWeights:
        cfg = self.cfg
        f32, bf = torch.float32, torch.bfloat16
        W = cfg.width
        kk = 3 * cfg.patch_size ** 2
        kpad = (kk + 63) // 64 * 64
        pw = torch.zeros(W, kpad, dtype=torch.float32)
        pw[:, :kk] = sd["visual.conv1.weight"].detach().float().cpu().reshape(W, kk)
        pos = sd["visual.positional_embedding"].detach().float().cpu()
        cls_pos0 = sd["visual.class_embedding"].detach().float().cpu() + pos[0]
        # closed form of the region pooling (textregion.py:183-195), folded in float64:
        #   r = ((mean @ Wv^T + bv) @ Wo^T + bo) @ proj = mean @ A + c
        ipw = sd["visual.attn_pool.attn.in_proj_weight"].detach().double().cpu()
        ipb = sd["visual.attn_pool.attn.in_proj_bias"].detach().double().cpu()
        Wv, bv = ipw[2 * W: 3 * W], ipb[2 * W: 3 * W]
        Wo = sd["visual.attn_pool.attn.out_proj.weight"].detach().double().cpu()
        bo = sd["visual.attn_pool.attn.out_proj.bias"].detach().double().cpu()
        proj = sd["visual.proj"].detach().double().cpu()
        A = Wv.T @ Wo.T @ proj                      # [W, D]
        cvec = (bv @ Wo.T + bo) @ proj              # [D]
        w = VitWeights()
        w.patch_w = ptr(self._dev(pw, bf)); w.patch_kpad = kpad
        w.cls_pos0 = ptr(self._dev(cls_pos0, f32)); w.pos = ptr(self._dev(pos, f32))
        w.ln_pre_w = ptr(self._dev(sd["visual.ln_pre.weight"], f32)); w.ln_pre_b = ptr(self._dev(sd["visual.ln_pre.bias"], f32))
        w.ln_post_w = ptr(self._dev(sd["visual.ln_post.weight"], f32)); w.ln_post_b = ptr(self._dev(sd["visual.ln_post.bias"], f32))
        w.blocks = self._blocks(sd, "visual.transformer.resblocks.", cfg.layers)
        w.pool_w = ptr(self._dev(A.T.float(), bf)); w.pool_b = ptr(self._dev(cvec.float(), f32))
        w.pool_b_empty = ptr(self._dev((bo @ proj).float(), f32))
        if self.has_text:
            w.tok_emb = ptr(self._dev(sd["token_embedding.weight"], f32))
            w.text_pos = ptr(self._dev(sd["positional_embedding"], f32))
            w.text_blocks = self._blocks(sd, "transformer.resblocks.", cfg.text_layers)
            w.ln_final_w = ptr(self._dev(sd["ln_final.weight"], f32)); w.ln_final_b = ptr(self._dev(sd["ln_final.bias"], f32))
            w.text_proj_w = ptr(self._dev(sd["text_projection"].detach().float().T, bf))
        return w

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ovo_encoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ calls
    def forward_features_from_pixels(self, pixels: torch.Tensor, n_layers: int = -1, ln_post: bool = True):
        """Test tap for E2: normalised pixels [n,3,S,S] f32 (device) -> tokens [n, seq, width] f32."""
        n = pixels.shape[0]
        pixels = pixels.to(self.device, torch.float32).contiguous()
        out = torch.empty(n, self.cfg.seq, self.cfg.width, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_encoder_load_pixels(self.handle, ptr(pixels), n, stream_ptr(self.device)), "load_pixels")
        check(self.lib.ovo_encoder_forward(self.handle, n, n_layers, int(ln_post), ptr(out), stream_ptr(self.device)), "forward")
        return out

    def forward_features(self, rgb_u8: torch.Tensor):
        """E1+E2: rgb uint8 [F,H,W,3] (device) -> tokens [F*n_img, seq, width]."""
        F_, H, W, _ = rgb_u8.shape
        per = C.c_int(0)
        check(self.lib.ovo_encoder_preprocess(self.handle, ptr(rgb_u8), F_, H, W, C.byref(per), stream_ptr(self.device)), "preprocess")
        n = F_ * per.value
        out = torch.empty(n, self.cfg.seq, self.cfg.width, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_encoder_forward(self.handle, n, -1, 1, ptr(out), stream_ptr(self.device)), "forward")
        return out

    def encode_regions(self, rgb_u8: torch.Tensor, masks: torch.Tensor, masks_per_frame=None) -> torch.Tensor:
        """CLIPGenerator.extract_clip, TextRegion branch.  rgb uint8 [H,W,3] or [F,H,W,3]; masks bool/uint8
        [M,H,W] (concatenated over frames) -> [M, output_dim] f32 unit-norm."""
        if rgb_u8.dim() == 3:
            rgb_u8 = rgb_u8[None]
        F_, H, W, _ = rgb_u8.shape
        rgb_u8 = rgb_u8.to(self.device, torch.uint8).contiguous()
        m8 = masks.to(self.device).to(torch.uint8).contiguous()
        M = m8.shape[0]
        counts = [M] if masks_per_frame is None else list(masks_per_frame)
        assert len(counts) == F_ and sum(counts) == M
        out = torch.empty(M, self.cfg.output_dim, device=self.device, dtype=torch.float32)
        if M == 0:
            return out
        arr = (C.c_int * F_)(*counts)
        check(self.lib.ovo_encode_regions(self.handle, ptr(rgb_u8), F_, H, W, ptr(m8), arr, ptr(out), stream_ptr(self.device)),
              "ovo_encode_regions")
        return out

    # ------------------------------------------------------------------ crop-based descriptors
    def install_pool_head(self, state_dict: dict | None = None, pool_heads: int = 8) -> None:
        """Packs `visual.attn_pool.*` + `visual.proj` (pe.py:44-87,540-541) for `encode_image`.  The probe is a
        parameter, so its query projection (pe.py:83-84) is folded here once, in float64."""
        sd = state_dict or self._sd_for_head
        if sd is None or "visual.attn_pool.probe" not in sd:
            raise RuntimeError("state_dict lacks visual.attn_pool.probe / layernorm / mlp: the crop-based embed types "
                               "need the full attention-pooling head")
        f32, bf = torch.float32, torch.bfloat16
        W = self.cfg.width
        p = "visual.attn_pool."
        ipw, ipb = sd[p + "attn.in_proj_weight"].detach().cpu(), sd[p + "attn.in_proj_bias"].detach().cpu()
        hd = W // pool_heads
        q = (sd[p + "probe"].detach().double().cpu().reshape(1, W) @ ipw[:W].double().T + ipb[:W].double()) * hd ** -0.5
        w = PoolHeadWeights()
        w.heads, w.mlp_width = pool_heads, sd[p + "mlp.c_fc.weight"].shape[0]
        w.q = ptr(self._dev(q.reshape(W).float(), f32))
        w.kv_w = ptr(self._dev(ipw[W:], bf)); w.kv_b = ptr(self._dev(ipb[W:], f32))
        w.out_w = ptr(self._dev(sd[p + "attn.out_proj.weight"], bf)); w.out_b = ptr(self._dev(sd[p + "attn.out_proj.bias"], f32))
        w.ln_w = ptr(self._dev(sd[p + "layernorm.weight"], f32)); w.ln_b = ptr(self._dev(sd[p + "layernorm.bias"], f32))
        w.fc_w = ptr(self._dev(sd[p + "mlp.c_fc.weight"], bf)); w.fc_b = ptr(self._dev(sd[p + "mlp.c_fc.bias"], f32))
        w.proj_w = ptr(self._dev(sd[p + "mlp.c_proj.weight"], bf)); w.proj_b = ptr(self._dev(sd[p + "mlp.c_proj.bias"], f32))
        w.vis_proj_w = ptr(self._dev(sd["visual.proj"].detach().float().T, bf))
        with torch.cuda.device(self.device):
            check(self.lib.ovo_encoder_set_pool_head(self.handle, C.byref(w)), "ovo_encoder_set_pool_head")
        self.has_pool_head = True
        self._sd_for_head = None

    def encode_images_from_pixels(self, pixels: torch.Tensor) -> torch.Tensor:
        """Test tap of pe.CLIP.encode_image: normalised pixels [n,3,S,S] -> [n, output_dim] (not normalised)."""
        pixels = pixels.to(self.device, torch.float32).contiguous()
        out = torch.empty(pixels.shape[0], self.cfg.output_dim, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_encode_images(self.handle, ptr(pixels), pixels.shape[0], ptr(out), stream_ptr(self.device)), "ovo_encode_images")
        return out

    def encode_crops(self, rgb_u8: torch.Tensor, masks: torch.Tensor, embed_type: str, mask_res: int = 384,
                     w_masked: float = 0.4418, w_global: float = 0.1, return_all: bool = False, bbox_margin: int = 50,
                     return_crops: bool = False):
        """CLIPGenerator.extract_clip, crop branch (clip_generator.py:136-158).  rgb uint8 [H,W,3]; masks bool/uint8
        [M,H,W] -> [M, output_dim] unit norm ([M,3,output_dim] with return_all)."""
        if not self.has_pool_head:
            raise RuntimeError("install_pool_head() first: the crop-based embed types use the attention-pooling head")
        H, W, _ = rgb_u8.shape
        rgb_u8 = rgb_u8.to(self.device, torch.uint8).contiguous()
        m8 = masks.to(self.device).to(torch.uint8).contiguous()
        M, D = m8.shape[0], self.cfg.output_dim
        vanilla = embed_type == "vanilla"
        out = torch.empty((M, 3, D) if (return_all and not vanilla) else (M, D), device=self.device, dtype=torch.float32)
        if M == 0:
            return out
        prm = CropParams(EMBED_TYPES[embed_type], int(return_all), mask_res, bbox_margin, w_masked, w_global)
        crops = None
        if return_crops:
            crops = torch.empty((M if vanilla else 2 * M), mask_res, mask_res, 3, device=self.device, dtype=torch.uint8)
        check(self.lib.ovo_encode_crops(self.handle, ptr(rgb_u8), H, W, ptr(m8), M, C.byref(prm), ptr(out), ptr(crops),
                                        stream_ptr(self.device)), "ovo_encode_crops")
        return (out, crops) if return_crops else out

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        """CLIP.encode_text: tokens int [T, ctx] -> [T, text_output_dim] f32 (not normalised)."""
        if not self.has_text:
            raise RuntimeError("encoder was built without a text tower")
        tok = tokens.to(self.device, torch.int32).contiguous()
        out = torch.empty(tok.shape[0], self.cfg.text_output_dim, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_encode_text(self.handle, ptr(tok), tok.shape[0], ptr(out), stream_ptr(self.device)), "ovo_encode_text")
        return out


def gemm_bf16(A: torch.Tensor, B: torch.Tensor, bias: torch.Tensor | None = None, force_bn: int = 0) -> torch.Tensor:
    """Test tap: C[M,N] f32 = A[M,K] bf16 . B[N,K]^T bf16 (+ bias)."""
    M, K = A.shape
    N = B.shape[0]
    out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    check(_lib.lib().ovo_gemm_bf16(ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, ptr(bias), ptr(out), N, force_bn,
                                   stream_ptr(A.device)), "ovo_gemm_bf16")
    return out
