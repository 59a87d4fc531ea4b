"""SAM-2 (image path) geometry and seeded random weights under the reference's own state_dict key names.

Geometry follows `Hiera.__init__` (thirdParty/segment-anything-2/sam2/modeling/backbones/hieradet.py:169-262) and
`configs/sam2.1/sam2.1_hiera_l.yaml`; the SAM heads follow `SAM2Base._build_sam_heads` (sam2_base.py:207-243)."""
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass
class BlockSpec:
    dim: int
    dim_out: int
    heads: int
    window: int        # 0 = global attention
    q_pool: bool       # MaxPool2d(2,2) on q and on the projected shortcut (stage transition)
    grid_in: int       # token grid side seen by the block input
    grid_out: int


@dataclass
class SamConfig:
    """SAM-2.1 Hiera-L defaults (sam2.1_hiera_l.yaml:9-28,88)."""
    image_size: int = 1024
    embed_dim: int = 144
    num_heads: int = 2
    stages: Tuple[int, ...] = (2, 6, 36, 4)
    global_att_blocks: Tuple[int, ...] = (23, 33, 43)
    window_spec: Tuple[int, ...] = (8, 4, 16, 8)
    pos_embed_bkg: Tuple[int, int] = (7, 7)
    d_model: int = 256               # neck / prompt / decoder width (fixed by the SAM heads)
    trunk_ln_eps: float = 1e-6       # hieradet.py:100
    # SAM heads (sam2_base.py:207-243)
    decoder_depth: int = 2
    decoder_heads: int = 8
    decoder_mlp: int = 2048
    num_mask_tokens: int = 4

    def blocks(self) -> List[BlockSpec]:
        """hieradet.py:205-243: the window size lags one block behind the stage change; q-pool blocks are the
        first block of stages 2..4."""
        stage_ends = [sum(self.stages[:i]) - 1 for i in range(1, len(self.stages) + 1)]
        q_pool_blocks = [x + 1 for x in stage_ends[:-1]]
        out = []
        dim, heads, cur_stage = self.embed_dim, self.num_heads, 1
        grid = self.image_size // 4
        for i in range(sum(self.stages)):
            dim_out = dim
            window = self.window_spec[cur_stage - 1]
            if i in self.global_att_blocks:
                window = 0
            if i - 1 in stage_ends:
                dim_out = dim * 2
                heads = heads * 2
                cur_stage += 1
            qp = i in q_pool_blocks
            out.append(BlockSpec(dim, dim_out, heads, window, qp, grid, grid // 2 if qp else grid))
            if qp:
                grid //= 2
            dim = dim_out
        return out

    def stage_ends(self) -> List[int]:
        return [sum(self.stages[:i]) - 1 for i in range(1, len(self.stages) + 1)]

    def channel_list(self) -> List[int]:
        b = self.blocks()
        return [b[i].dim_out for i in self.stage_ends()[::-1]]


def tiny_sam_config() -> SamConfig:
    """Same code paths as Hiera-L (head_dim 72, windows 8/4/16/8, one global block, three q-pool transitions,
    1024^2 input) at ~1/300 of the work: the golden fixtures and the CPU oracle stay small."""
    return SamConfig(embed_dim=72, num_heads=1, stages=(1, 2, 3, 1), global_att_blocks=(4,))


def random_state_dict(cfg: SamConfig, seed: int = 0):
    """Seeded random SAM-2 weights (image path only) with the reference's key names.  No checkpoint is available
    offline; scales keep activations O(1) so that parity tolerances mean something."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}

    def linear(name, out_f, in_f, wstd=None):
        sd[name + ".weight"] = rn(out_f, in_f, std=wstd if wstd is not None else in_f ** -0.5)
        sd[name + ".bias"] = rn(out_f, std=0.02)

    def norm(name, n):
        sd[name + ".weight"] = 1 + rn(n, std=0.05)
        sd[name + ".bias"] = rn(n, std=0.02)

    t = "image_encoder.trunk."
    sd[t + "patch_embed.proj.weight"] = rn(cfg.embed_dim, 3, 7, 7, std=147 ** -0.5)
    sd[t + "patch_embed.proj.bias"] = rn(cfg.embed_dim, std=0.02)
    sd[t + "pos_embed"] = rn(1, cfg.embed_dim, *cfg.pos_embed_bkg, std=0.2)
    sd[t + "pos_embed_window"] = rn(1, cfg.embed_dim, cfg.window_spec[0], cfg.window_spec[0], std=0.2)
    for i, b in enumerate(cfg.blocks()):
        p = f"{t}blocks.{i}."
        norm(p + "norm1", b.dim)
        linear(p + "attn.qkv", 3 * b.dim_out, b.dim)
        linear(p + "attn.proj", b.dim_out, b.dim_out, wstd=0.5 * b.dim_out ** -0.5)
        norm(p + "norm2", b.dim_out)
        linear(p + "mlp.layers.0", 4 * b.dim_out, b.dim_out)
        linear(p + "mlp.layers.1", b.dim_out, 4 * b.dim_out, wstd=0.5 * (4 * b.dim_out) ** -0.5)
        if b.dim != b.dim_out:
            linear(p + "proj", b.dim_out, b.dim)
    D = cfg.d_model
    for j, c in enumerate(cfg.channel_list()):
        sd[f"image_encoder.neck.convs.{j}.conv.weight"] = rn(D, c, 1, 1, std=c ** -0.5)
        sd[f"image_encoder.neck.convs.{j}.conv.bias"] = rn(D, std=0.02)
    sd["no_mem_embed"] = rn(1, 1, D, std=0.02)
    pe = "sam_prompt_encoder."
    sd[pe + "pe_layer.positional_encoding_gaussian_matrix"] = rn(2, D // 2)
    for i in range(4):
        sd[pe + f"point_embeddings.{i}.weight"] = rn(1, D, std=0.5)
    sd[pe + "not_a_point_embed.weight"] = rn(1, D, std=0.5)
    sd[pe + "no_mask_embed.weight"] = rn(1, D, std=0.5)
    md = "sam_mask_decoder."

    def attention(name, internal):
        linear(name + ".q_proj", internal, D)
        linear(name + ".k_proj", internal, D)
        linear(name + ".v_proj", internal, D)
        linear(name + ".out_proj", D, internal)

    for l in range(cfg.decoder_depth):
        p = f"{md}transformer.layers.{l}."
        attention(p + "self_attn", D)
        attention(p + "cross_attn_token_to_image", D // 2)
        attention(p + "cross_attn_image_to_token", D // 2)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            norm(p + n, D)
        linear(p + "mlp.layers.0", cfg.decoder_mlp, D)
        linear(p + "mlp.layers.1", D, cfg.decoder_mlp)
    attention(md + "transformer.final_attn_token_to_image", D // 2)
    norm(md + "transformer.norm_final_attn", D)
    sd[md + "iou_token.weight"] = rn(1, D, std=0.5)
    sd[md + "mask_tokens.weight"] = rn(cfg.num_mask_tokens, D, std=0.5)
    sd[md + "obj_score_token.weight"] = rn(1, D, std=0.5)
    sd[md + "output_upscaling.0.weight"] = rn(D, D // 4, 2, 2, std=D ** -0.5)     # ConvTranspose2d [in, out, kh, kw]
    sd[md + "output_upscaling.0.bias"] = rn(D // 4, std=0.02)
    norm(md + "output_upscaling.1", D // 4)
    sd[md + "output_upscaling.3.weight"] = rn(D // 4, D // 8, 2, 2, std=(D // 4) ** -0.5)
    sd[md + "output_upscaling.3.bias"] = rn(D // 8, std=0.02)
    sd[md + "conv_s0.weight"] = rn(D // 8, D, 1, 1, std=D ** -0.5); sd[md + "conv_s0.bias"] = rn(D // 8, std=0.02)
    sd[md + "conv_s1.weight"] = rn(D // 4, D, 1, 1, std=D ** -0.5); sd[md + "conv_s1.bias"] = rn(D // 4, std=0.02)
    for i in range(cfg.num_mask_tokens):
        linear(f"{md}output_hypernetworks_mlps.{i}.layers.0", D, D)
        linear(f"{md}output_hypernetworks_mlps.{i}.layers.1", D, D)
        linear(f"{md}output_hypernetworks_mlps.{i}.layers.2", D // 8, D)
    linear(md + "iou_prediction_head.layers.0", 256, D)
    linear(md + "iou_prediction_head.layers.1", 256, 256)
    linear(md + "iou_prediction_head.layers.2", cfg.num_mask_tokens, 256)
    linear(md + "pred_obj_score_head.layers.0", D, D)
    linear(md + "pred_obj_score_head.layers.1", D, D)
    linear(md + "pred_obj_score_head.layers.2", 1, D)
    return sd
