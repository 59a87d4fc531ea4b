"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's CROP-BASED descriptor path (SURVEY §8f rank 2,
row E6 crop branch): embed types `vanilla`, `fixed_weights`, `hovsg`, `adaptive_weights`, `concept_fusion`
(and `return_all`), restated on the vendored Perception-Encoder CLIP (`pe.CLIP.encode_image`).

NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import it.

Reference lines restated (paths relative to /root/reference):
  cg = ovo/entities/clip_generator.py     (extract_clip :125-158, encode_image :111-122)
  su = ovo/utils/segment_utils.py         (segmap2segimg :29-41, batched_mask_to_box :43-96, xyxy->xywh :99-104,
                                           seg_img_from_image :128-136, get_seg_img :138-142, get_bbox_img :144-147,
                                           pad_img :149-157, increase_bbox_by_margin :159-182)
  cu = ovo/utils/clip_utils.py            (fuse_clips :21-48, siglip_cosine_similarity :10-14)
  pe = thirdParty/perception_models/core/vision_encoder/pe.py (AttentionPooling :44-87, forward :535-543, _pool :486-497)
  torchvision F.resize on a uint8 tensor (un-vendored dependency; the installed torchvision is the oracle):
  float32 anti-aliased bilinear interpolation, torch.round (half to even), cast back to uint8.

The reference's own crop branch loads its encoder through open_clip (cu:51-88), which is not vendored; the PE card it
lists there (`PE-Core-L-14-336`) is the architecture of the vendored `pe.CLIP`, so the goldens
(tests/golden/crops.npz, oracle/gen_golden.py:gen_crops) run the UNMODIFIED `CLIPGenerator.extract_clip`,
`segmap2segimg` and `fuse_clips` with `load_clip_model` returning the vendored `pe.CLIP` + its own transform.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import encoder as OE

EMBED_TYPES = ("vanilla", "fixed_weights", "hovsg", "adaptive_weights", "concept_fusion")


# ----------------------------------------------------------------------------------------------
# segmap2segimg (su:29-182)
# ----------------------------------------------------------------------------------------------
def mask_boxes_xywh(masks: np.ndarray) -> np.ndarray:
    """[M,H,W] bool -> [M,4] int64 (x, y, w, h).  su:43-104: edges are the min / max set row / column INDEX, so
    w = right - left and h = bottom - top (one less than the pixel extent; kept, not "fixed"); an empty mask
    gives [0,0,0,0]."""
    M, H, W = masks.shape
    out = np.zeros((M, 4), np.int64)
    for i in range(M):
        rows = np.flatnonzero(masks[i].any(axis=1))
        cols = np.flatnonzero(masks[i].any(axis=0))
        if len(rows) == 0:
            continue
        out[i] = (cols[0], rows[0], cols[-1] - cols[0], rows[-1] - rows[0])
    return out


def resize_u8(img_u8: torch.Tensor, size: int) -> torch.Tensor:
    """torchvision F.resize(img uint8 [C,h,w], (size,size)): f32 AA bilinear, round half-even, uint8."""
    if img_u8.shape[1] == 0 or img_u8.shape[2] == 0:
        raise RuntimeError("Input and output sizes should be greater than 0 (empty crop)")
    return torch.round(OE.aa_resize(img_u8.float(), size, size)).to(torch.uint8)


def margin_box(x, y, w, h, margin):
    """su:159-182 (only the left / top edges are clamped; slicing clamps the others)."""
    x, y, w, h = x - margin, y - margin, w + 2 * margin, h + 2 * margin
    if x < 0:
        w += x
        x = 0
    if y < 0:
        h += y
        y = 0
    return x, y, w, h


def seg_images(masks: np.ndarray, image_u8_chw: torch.Tensor, also_bbox: bool, margin: int = 50, out_l: int = 224):
    """su:29-41,128-157 -> uint8 [M, 6 if also_bbox else 3, out_l, out_l]: channels 0-2 = the masked crop (zero outside
    the mask; squashed when also_bbox, zero-padded to a centred square otherwise), channels 3-5 = the margin crop."""
    boxes = mask_boxes_xywh(masks)
    out = []
    for i in range(masks.shape[0]):
        x, y, w, h = (int(v) for v in boxes[i])
        m = torch.from_numpy(masks[i, y:y + h, x:x + w])
        seg = image_u8_chw[:, y:y + h, x:x + w] * m[None].to(torch.uint8)
        if also_bbox:
            bx, by, bw, bh = margin_box(x, y, w, h, margin)
            bbox = image_u8_chw[:, by:by + bh, bx:bx + bw]
            out.append(torch.cat([resize_u8(seg, out_l), resize_u8(bbox, out_l)], 0))
        else:
            side = max(w, h)
            pad = torch.zeros(3, side, side, dtype=torch.uint8)
            if h > w:
                pad[:, :, (h - w) // 2:(h - w) // 2 + w] = seg
            else:
                pad[:, (w - h) // 2:(w - h) // 2 + h, :] = seg
            out.append(resize_u8(pad, out_l))
    return torch.stack(out) if out else torch.zeros(0, 6 if also_bbox else 3, out_l, out_l, dtype=torch.uint8)


# ----------------------------------------------------------------------------------------------
# encode_image (cg:111-122 -> pe:535-543): Resize((S,S)) + Normalize, ViT, attention pooling, projection
# ----------------------------------------------------------------------------------------------
def attn_pool(tokens: torch.Tensor, W: dict, heads: int, eps: float = 1e-5) -> torch.Tensor:
    """pe:44-87: one learned probe attends to all tokens (nn.MultiheadAttention), then x + mlp(layernorm(x)).
    tokens [n, S, width] -> [n, width]."""
    n, S, D = tokens.shape
    hd = D // heads
    p = "visual.attn_pool."
    ipw, ipb = W[p + "attn.in_proj_weight"], W[p + "attn.in_proj_bias"]
    q = F.linear(W[p + "probe"].reshape(1, D), ipw[:D], ipb[:D])                # [1, D]
    k = F.linear(tokens, ipw[D:2 * D], ipb[D:2 * D]).view(n, S, heads, hd)
    v = F.linear(tokens, ipw[2 * D:], ipb[2 * D:]).view(n, S, heads, hd)
    s = torch.einsum("hd,nshd->nhs", q.view(heads, hd), k) * (hd ** -0.5)
    a = torch.einsum("nhs,nshd->nhd", torch.softmax(s, dim=-1), v).reshape(n, D)
    x = F.linear(a, W[p + "attn.out_proj.weight"], W[p + "attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), W[p + "layernorm.weight"], W[p + "layernorm.bias"], eps)
    h = F.gelu(F.linear(h, W[p + "mlp.c_fc.weight"], W[p + "mlp.c_fc.bias"]))
    return x + F.linear(h, W[p + "mlp.c_proj.weight"], W[p + "mlp.c_proj.bias"])


def encode_image(pixels01: torch.Tensor, W: dict, cfg: OE.VitCfg, pool_heads: int) -> torch.Tensor:
    """pixels01 [n,3,h,w] f32 in [0,1] -> [n, output_dim] (NOT normalised)."""
    S = cfg.image_size
    px = torch.stack([(OE.aa_resize(p, S, S) - 0.5) / 0.5 for p in pixels01])
    tok = OE.vit_forward_features(px, W, cfg, norm=True)
    return attn_pool(tok, W, pool_heads) @ W["visual.proj"]


# ----------------------------------------------------------------------------------------------
# fuse_clips (cu:21-48)
# ----------------------------------------------------------------------------------------------
def _cos(a, b, eps=1e-6):
    return F.cosine_similarity(a, b, dim=-1, eps=eps)


def fuse_clips(clip_g, clip_seg, clip_bbox, embed_type: str, w_masked: float, w_global: float) -> torch.Tensor:
    nrm = lambda t: F.normalize(t, p=2, dim=-1)
    if embed_type in ("hovsg", "fixed_weights"):
        clip_l = nrm(clip_seg * w_masked + clip_bbox * (1 - w_masked))
        if embed_type == "fixed_weights":
            wg = w_global
        else:
            wg = torch.softmax(_cos(clip_g, clip_l), dim=0).unsqueeze(1)          # softmax ACROSS the masks of the frame
        return nrm(clip_g * wg + clip_l * (1 - wg))
    if embed_type == "adaptive_weights":
        wl = (_cos(clip_seg, clip_bbox) * w_masked).unsqueeze(-1)
        clip_l = nrm(clip_seg * wl + clip_bbox * (1 - wl))
        wg = (_cos(clip_g, clip_l) * w_global).unsqueeze(-1)
        return nrm(clip_g * wg + clip_l * (1 - wg))
    if embed_type == "concept_fusion":
        wg = torch.softmax(_cos(clip_g, clip_bbox), dim=0).unsqueeze(-1)
        return nrm(wg * clip_g + (1 - wg) * clip_bbox)
    return clip_seg


def extract_clip(image_u8_hwc: np.ndarray, masks: np.ndarray, W: dict, cfg: OE.VitCfg, embed_type: str,
                 mask_res: int = 384, w_masked: float = 0.4418, w_global: float = 0.1, return_all: bool = False,
                 pool_heads: int = 8, margin: int = 50) -> torch.Tensor:
    """cg:125-158, crop branch -> [M, D] ([M,3,D] with return_all)."""
    img = torch.from_numpy(np.ascontiguousarray(image_u8_hwc.transpose(2, 0, 1)))
    nrm = lambda t: F.normalize(t, p=2, dim=-1)
    if masks.shape[0] == 0:
        return torch.zeros(0)
    if embed_type == "vanilla":
        seg = seg_images(masks, img, False, margin, mask_res).float() / 255.0
        return nrm(encode_image(seg[:, :3], W, cfg, pool_heads))
    clip_g = nrm(encode_image(img[None].float() / 255.0, W, cfg, pool_heads))
    seg = seg_images(masks, img, True, margin, mask_res).float() / 255.0
    n = seg.shape[0]
    e = nrm(encode_image(torch.cat([seg[:, :3], seg[:, 3:]]), W, cfg, pool_heads))
    g = clip_g.repeat(n, 1)
    if return_all:
        return torch.stack([g, e[:n], e[n:]], dim=1)
    return fuse_clips(g, e[:n], e[n:], embed_type, w_masked, w_global)


def siglip_similarity(txt: torch.Tensor, img: torch.Tensor, logit_scale: float, logit_bias: float) -> torch.Tensor:
    """cu:10-14."""
    return torch.sigmoid(img @ txt.T * float(np.exp(logit_scale)) + logit_bias)


# ----------------------------------------------------------------------------------------------
# embed_type `learned`: WeightsPredictorMerger (ovo/entities/clips_merging.py:26-56)
# ----------------------------------------------------------------------------------------------
def weights_predictor_merge(input_clips: torch.Tensor, W: dict, nhead: int = 8, eps: float = 1e-5) -> torch.Tensor:
    """input_clips [B, 3, D] (global, masked crop, margin crop) -> merged [B, D], unit norm.
    `att_encoder` = nn.TransformerEncoder of post-norm nn.TransformerEncoderLayer(activation relu) (cm:29-36; eval mode, no dropout);
    `mlp` = Linear, LeakyReLU, n x (Linear, LeakyReLU), Linear (cm:13-24); the soft-max runs over the three clips, per channel when
    the MLP emits 3*D weights, per clip when it emits 3 (cm:48-53).  W uses the module's own state_dict keys."""
    b, n, d = input_clips.shape
    hd = d // nhead
    x = input_clips
    L = 0
    while f"att_encoder.layers.{L}.self_attn.in_proj_weight" in W:
        p = f"att_encoder.layers.{L}."
        qkv = F.linear(x, W[p + "self_attn.in_proj_weight"], W[p + "self_attn.in_proj_bias"])
        q, k, v = (t.view(b, n, nhead, hd).transpose(1, 2) for t in qkv.split(d, dim=-1))
        a = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1) @ v
        a = F.linear(a.transpose(1, 2).reshape(b, n, d), W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + a, (d,), W[p + "norm1.weight"], W[p + "norm1.bias"], eps)
        f = F.linear(F.relu(F.linear(x, W[p + "linear1.weight"], W[p + "linear1.bias"])), W[p + "linear2.weight"], W[p + "linear2.bias"])
        x = F.layer_norm(x + f, (d,), W[p + "norm2.weight"], W[p + "norm2.bias"], eps)
        L += 1
    h = x.flatten(-2, -1)
    idx = sorted(int(k.split(".")[1]) for k in W if k.startswith("mlp.") and k.endswith(".weight"))
    for j, i in enumerate(idx):
        h = F.linear(h, W[f"mlp.{i}.weight"], W[f"mlp.{i}.bias"])
        if j + 1 < len(idx):
            h = F.leaky_relu(h, 0.01)
    if h.shape[-1] != 3:
        w = torch.softmax(h.reshape(b, n, d), dim=-2)
    else:
        w = torch.softmax(h, dim=-1).unsqueeze(-1)
    return F.normalize((input_clips * w).sum(-2), dim=-1)
