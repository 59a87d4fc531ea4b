"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's image/text encoder path.

This is the oracle the CUDA path is checked against.  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Every function cites the reference lines it restates (paths relative to /root/reference):
  tr  = ovo/entities/textregion.py
  pe  = thirdParty/perception_models/core/vision_encoder/pe.py
  rp  = thirdParty/perception_models/core/vision_encoder/rope.py
  tf  = thirdParty/perception_models/core/vision_encoder/transforms.py
Pinned against the reference itself (imported in the build container) by tests/golden/*.npz, generated
with oracle/gen_golden.py; the reference ships no golden vectors of its own for this path.

Weights are passed as a plain dict with the reference's own state_dict key names
(`visual.transformer.resblocks.{i}.attn.in_proj_weight`, ...), see SURVEY Appendix B.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class VitCfg:
    image_size: int = 336
    patch_size: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp_width: int = 4096
    output_dim: int = 1024
    ln_eps: float = 1e-5
    # text tower
    text_ctx: int = 32
    text_width: int = 1024
    text_heads: int = 16
    text_layers: int = 24
    text_mlp_width: int = 4096
    vocab_size: int = 49408

    @property
    def grid(self):
        return self.image_size // self.patch_size

    @property
    def seq(self):
        return self.grid * self.grid + 1


# ----------------------------------------------------------------------------------------------
# E1: crops + anti-aliased bilinear resize + normalize   (tr:104-134, tf:19-26)
# ----------------------------------------------------------------------------------------------
def aa_weights(n_in: int, n_out: int):
    """Per-axis triangle-filter weights of torchvision Resize(bilinear, antialias=True)
    (ATen _upsample_bilinear2d_aa).  Returns (xmin[n_out] int, size[n_out] int, w[n_out, kmax] f32).
    SURVEY Appendix A1."""
    scale = n_in / n_out
    support = max(scale, 1.0)
    invscale = 1.0 / max(scale, 1.0)
    kmax = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(n_out, np.int32)
    size = np.zeros(n_out, np.int32)
    w = np.zeros((n_out, kmax), np.float32)
    for i in range(n_out):
        center = scale * (i + 0.5)
        lo = max(0, int(center - support + 0.5))
        hi = min(n_in, int(center + support + 0.5))
        ws = np.array([max(0.0, 1.0 - abs((j - center + 0.5) * invscale)) for j in range(lo, hi)], np.float64)
        ws = ws / ws.sum()
        xmin[i], size[i] = lo, hi - lo
        w[i, : hi - lo] = ws.astype(np.float32)
    return xmin, size, w


def aa_resize(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """img [C,h,w] f32 -> [C,out_h,out_w]; separable, horizontal pass first like ATen."""
    C, h, w = img.shape
    if (h, w) == (out_h, out_w):
        return img.clone()
    xmin, xs, xw = aa_weights(w, out_w)
    ymin, ys, yw = aa_weights(h, out_h)
    Wx = torch.zeros(out_w, w)
    for i in range(out_w):
        Wx[i, xmin[i]: xmin[i] + xs[i]] = torch.from_numpy(xw[i, : xs[i]])
    Wy = torch.zeros(out_h, h)
    for i in range(out_h):
        Wy[i, ymin[i]: ymin[i] + ys[i]] = torch.from_numpy(yw[i, : ys[i]])
    tmp = torch.einsum("chw,ow->cho", img, Wx)
    return torch.einsum("cho,ph->cpo", tmp, Wy)


def crop_boxes(h: int, w: int, crop_size: int):
    """Tile boxes (y1,y2,x1,x2) of the multi_resolution strategy (tr:114-128)."""
    nh, nw = max(h // crop_size, 1), max(w // crop_size, 1)
    ch, cw = int(np.ceil(h / nh)), int(np.ceil(w / nw))
    boxes = []
    for hi in range(nh):
        for wi in range(nw):
            y1, x1 = hi * ch, wi * cw
            y2, x2 = min(y1 + ch, h), min(x1 + cw, w)
            y1, x1 = max(y2 - ch, 0), max(x2 - cw, 0)
            boxes.append((y1, y2, x1, x2))
    return nh, nw, boxes


def preprocess(image01: torch.Tensor, cfg: VitCfg) -> torch.Tensor:
    """image01 [3,H,W] f32 in [0,1] -> [n_img,3,S,S] normalised with mean=std=0.5 (tr:104-134)."""
    S = cfg.image_size
    _, H, W = image01.shape
    _, _, boxes = crop_boxes(H, W, S)
    outs = [(aa_resize(image01, S, S) - 0.5) / 0.5]
    for (y1, y2, x1, x2) in boxes:
        outs.append((aa_resize(image01[:, y1:y2, x1:x2], S, S) - 0.5) / 0.5)
    return torch.stack(outs)


# ----------------------------------------------------------------------------------------------
# E2: ViT forward_features(norm=True)   (pe:499-533, 216-225, 123-150; rp:315-347, 40-62)
# ----------------------------------------------------------------------------------------------
def rope_table(grid: int, head_dim: int):
    """cos/sin [1+grid*grid, head_dim/2] for PE's 2D RoPE with a cls token (rp:315-340).
    Pair i<hd/4 rotates by (x+1)*theta_i, pair i>=hd/4 by (y+1)*theta_{i-hd/4};
    theta_j = 10000^(-2j/(hd/2)); cls row = angle 0."""
    half = head_dim // 2          # rotary dim handed to RotaryEmbedding (rp:313 dim//2)
    nf = half // 2                # number of distinct freqs per axis
    theta = 1.0 / (10000 ** (torch.arange(0, half, 2)[:nf].float() / half))
    ys, xs = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing="ij")
    ax = (xs.reshape(-1, 1).float() + 1) * theta[None]     # [g*g, nf]
    ay = (ys.reshape(-1, 1).float() + 1) * theta[None]
    ang = torch.cat([ax, ay], dim=1)                        # [g*g, hd/2] one angle per pair
    ang = torch.cat([torch.zeros(1, ang.shape[1]), ang], dim=0)
    return ang.cos(), ang.sin()


def apply_rope(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """t [..., S, hd]; interleaved pairs (2i,2i+1): out[2i]=a*cos-b*sin, out[2i+1]=b*cos+a*sin (rp:32-62)."""
    a, b = t[..., 0::2], t[..., 1::2]
    o = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1)
    return o.flatten(-2)


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def resblock(x, W, pfx, heads, eps, rope=None, causal=False):
    """One ResidualAttentionBlock (pe:216-225).  Vision blocks: SelfAttention with RoPE (pe:123-150);
    text blocks: nn.MultiheadAttention with the causal additive mask (pe:167-170, 621-627)."""
    B, S, D = x.shape
    hd = D // heads
    h = _ln(x, W[pfx + "ln_1.weight"], W[pfx + "ln_1.bias"], eps)
    qkv = F.linear(h, W[pfx + "attn.in_proj_weight"], W[pfx + "attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, heads, hd).transpose(1, 2)
    k = k.view(B, S, heads, hd).transpose(1, 2)
    v = v.view(B, S, heads, hd).transpose(1, 2)
    if rope is not None:
        q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    if causal:
        s = s + torch.full((S, S), float("-inf")).triu(1)
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, S, D)
    x = x + F.linear(a, W[pfx + "attn.out_proj.weight"], W[pfx + "attn.out_proj.bias"])
    h = _ln(x, W[pfx + "ln_2.weight"], W[pfx + "ln_2.bias"], eps)
    h = F.gelu(F.linear(h, W[pfx + "mlp.c_fc.weight"], W[pfx + "mlp.c_fc.bias"]))
    x = x + F.linear(h, W[pfx + "mlp.c_proj.weight"], W[pfx + "mlp.c_proj.bias"])
    return x


def vit_forward_features(pixels: torch.Tensor, W: dict, cfg: VitCfg, n_layers: int = -1,
                         norm: bool = True, taps: list | None = None) -> torch.Tensor:
    """pixels [n,3,S,S] normalised -> tokens [n, 1+g*g, width] (pe:499-533, image at native 336 so the
    abs pos-emb needs no interpolation, pe:465-466)."""
    n = pixels.shape[0]
    p, Wd = cfg.patch_size, cfg.width
    x = F.conv2d(pixels, W["visual.conv1.weight"], stride=p)           # pe:509
    x = x.permute(0, 2, 3, 1).reshape(n, -1, Wd)
    cls = W["visual.class_embedding"].view(1, 1, -1).expand(n, -1, -1)
    x = torch.cat([cls, x], dim=1) + W["visual.positional_embedding"][None]
    x = _ln(x, W["visual.ln_pre.weight"], W["visual.ln_pre.bias"], cfg.ln_eps)
    if taps is not None:
        taps.append(x.clone())
    rope = rope_table(cfg.grid, Wd // cfg.heads)
    L = cfg.layers if n_layers < 0 else n_layers
    for i in range(L):
        x = resblock(x, W, f"visual.transformer.resblocks.{i}.", cfg.heads, cfg.ln_eps, rope=rope)
        if taps is not None:
            taps.append(x.clone())
    if norm:
        x = _ln(x, W["visual.ln_post.weight"], W["visual.ln_post.bias"], cfg.ln_eps)
    return x


# ----------------------------------------------------------------------------------------------
# E3: multi-resolution token canvas   (tr:9-28)
# ----------------------------------------------------------------------------------------------
def bilinear_src(n_in: int, n_out: int):
    """F.interpolate(mode='bilinear', align_corners=False) taps: (i0, i1, lambda1) per output index."""
    scale = n_in / n_out
    i0 = np.zeros(n_out, np.int64); i1 = np.zeros(n_out, np.int64); l1 = np.zeros(n_out, np.float32)
    for o in range(n_out):
        src = max(scale * (o + 0.5) - 0.5, 0.0)
        a = int(math.floor(src))
        a = min(a, n_in - 1)
        b = min(a + 1, n_in - 1)
        i0[o], i1[o], l1[o] = a, b, np.float32(src - a)
    return i0, i1, l1


def token_canvas(tokens: torch.Tensor, cfg: VitCfg, nh: int, nw: int) -> torch.Tensor:
    """tokens [n_img, g*g, D] (cls stripped, tr:165-166) -> canvas [nh*g*nw*g, D]  (tr:9-28):
    global tokens bilinearly up-sampled to [nh*g, nw*g], then 0.5*up + crop tokens tile by tile."""
    g, D = cfg.grid, tokens.shape[-1]
    ph, pw = nh * g, nw * g
    glob = tokens[0].view(g, g, D)
    y0, y1, ly = bilinear_src(g, ph)
    x0, x1, lx = bilinear_src(g, pw)
    ly_t = torch.from_numpy(ly).view(ph, 1, 1); lx_t = torch.from_numpy(lx).view(1, pw, 1)
    top = glob[y0][:, x0] * (1 - lx_t) + glob[y0][:, x1] * lx_t
    bot = glob[y1][:, x0] * (1 - lx_t) + glob[y1][:, x1] * lx_t
    up = top * (1 - ly_t) + bot * ly_t                                   # [ph,pw,D]
    canvas = up.clone()
    cid = 1
    for hi in range(nh):
        for wi in range(nw):
            ys, xs = hi * g, wi * g
            canvas[ys:ys + g, xs:xs + g] = 0.5 * up[ys:ys + g, xs:xs + g] + tokens[cid].view(g, g, D)
            cid += 1
    return canvas.reshape(ph * pw, D)


# ----------------------------------------------------------------------------------------------
# E4: region masks -> token-grid masks   (tr:145-161, used as `<= 0` at tr:187)
# ----------------------------------------------------------------------------------------------
def feature_masks(masks: np.ndarray, ph: int, pw: int) -> np.ndarray:
    """masks [M,H,W] bool -> [M, ph*pw] bool: token belongs to the mask iff any bilinear tap with a
    non-zero weight is set."""
    M, H, W = masks.shape
    y0, y1, ly = bilinear_src(H, ph)
    x0, x1, lx = bilinear_src(W, pw)
    m = masks.astype(bool)
    out = np.zeros((M, ph, pw), bool)
    wy = [(y0, 1 - ly), (y1, ly)]
    wx = [(x0, 1 - lx), (x1, lx)]
    for yi, yw in wy:
        for xi, xw in wx:
            nz = (yw[:, None] > 0) & (xw[None, :] > 0)
            out |= m[:, yi][:, :, xi] & nz[None]
    return out.reshape(M, ph * pw)


# ----------------------------------------------------------------------------------------------
# E5: mask-restricted attention pooling -> projection -> L2 norm   (tr:163-195, pe:44-87)
# ----------------------------------------------------------------------------------------------
def region_pool(canvas: torch.Tensor, fmask: torch.Tensor, W: dict, cfg: VitCfg) -> torch.Tensor:
    """Closed form of tr:183-195: all keys are identical (tr:185-186), so the softmax over
    un-padded keys is uniform and every head sees the same weights:
        out = normalize( out_proj( W_v . mean_{p in mask} x_p + b_v ) @ proj ).
    A mask with no token has every key padded; torch's MHA (safe softmax, torch >= 2.5: verified against the
    reference in tests/golden/encoder_tiny.npz) then attends to nothing, the attention output is 0 and the
    region feature is normalize(out_proj.bias @ proj)."""
    D = cfg.width
    Wv = W["visual.attn_pool.attn.in_proj_weight"][2 * D: 3 * D]
    bv = W["visual.attn_pool.attn.in_proj_bias"][2 * D: 3 * D]
    fm = fmask.float()
    cnt = fm.sum(dim=1, keepdim=True)
    mean = (fm @ canvas) / cnt.clamp_min(1.0)
    v = F.linear(mean, Wv, bv)
    v = torch.where(cnt > 0, v, torch.zeros_like(v))
    o = F.linear(v, W["visual.attn_pool.attn.out_proj.weight"], W["visual.attn_pool.attn.out_proj.bias"])
    r = o @ W["visual.proj"]
    return r / r.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def encode_regions(image_u8_hwc: np.ndarray, masks: np.ndarray, W: dict, cfg: VitCfg) -> torch.Tensor:
    """CLIPGenerator.extract_clip, TextRegion branch (clip_generator.py:134-135 -> tr:197-203).
    image [H,W,3] uint8, masks [M,H,W] bool -> [M, output_dim] unit-norm f32."""
    img = torch.from_numpy(np.ascontiguousarray(image_u8_hwc.transpose(2, 0, 1))).float() / 255.0
    H, Wd = img.shape[1:]
    nh, nw, _ = crop_boxes(H, Wd, cfg.image_size)
    px = preprocess(img, cfg)
    tok = vit_forward_features(px, W, cfg)[:, 1:]
    canvas = token_canvas(tok, cfg, nh, nw)
    fm = torch.from_numpy(feature_masks(masks, nh * cfg.grid, nw * cfg.grid))
    return region_pool(canvas, fm, W, cfg)


# ----------------------------------------------------------------------------------------------
# Q1: text tower   (pe:671-695; clip_generator.py:161-199)
# ----------------------------------------------------------------------------------------------
def text_forward(tokens: torch.Tensor, W: dict, cfg: VitCfg) -> torch.Tensor:
    """tokens [T,ctx] int -> [T, output_dim] (un-normalised), pooled at argmax(token id) = EOT (pe:662-665)."""
    x = W["token_embedding.weight"][tokens] + W["positional_embedding"][: tokens.shape[1]]
    for i in range(cfg.text_layers):
        x = resblock(x, W, f"transformer.resblocks.{i}.", cfg.text_heads, cfg.ln_eps, causal=True)
    x = _ln(x, W["ln_final.weight"], W["ln_final.bias"], cfg.ln_eps)
    pooled = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)]
    return pooled @ W["text_projection"]


def text_bank(per_query_tokens: list, W: dict, cfg: VitCfg) -> torch.Tensor:
    """get_embed_txt_similarity's embedding rule (clip_generator.py:191-196): per query
    normalize(mean_over_templates(normalize(encode_text(tokens))))."""
    rows = []
    for tok in per_query_tokens:
        e = text_forward(tok, W, cfg)
        e = e / e.norm(dim=-1, keepdim=True)
        rows.append(F.normalize(e.mean(0, keepdim=True), p=2, dim=-1)[0])
    return torch.stack(rows)


def cosine_query(bank: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
    """clip_cosine_similarity (clip_utils.py:16-19): [I,D] x [Q,D]^T -> [I,Q]; no logit scale for PE cards."""
    return bank @ text.to(bank.dtype).T


def cfg_from_reference(vcfg, tcfg) -> VitCfg:
    return VitCfg(image_size=vcfg.image_size, patch_size=vcfg.patch_size, width=vcfg.width, layers=vcfg.layers,
                  heads=vcfg.heads, mlp_width=int(vcfg.width * vcfg.mlp_ratio), output_dim=vcfg.output_dim,
                  text_ctx=tcfg.context_length, text_width=tcfg.width, text_heads=tcfg.heads,
                  text_layers=tcfg.layers, text_mlp_width=int(tcfg.width * tcfg.mlp_ratio),
                  vocab_size=tcfg.vocab_size)
