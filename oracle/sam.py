"""TEST INFRASTRUCTURE — CPU restatement (torch f32 / numpy) of the reference's SAM-2 image path:
image transform, Hiera trunk, FPN neck, prompt encoder, two-way mask decoder and the automatic-mask-generator
post-processing, working directly from a state_dict with the reference's key names.

Pinned against the UNMODIFIED reference (SAM2Base / SAM2ImagePredictor / SAM2AutomaticMaskGenerator run in the build
container) by oracle/gen_golden.py gen_sam -> tests/golden/sam_tiny.npz and tests/test_oracle_golden.py.
Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Paths cited below are relative to thirdParty/segment-anything-2/sam2/ of the reference."""
import math

import numpy as np
import torch
import torch.nn.functional as F

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


# ----------------------------------------------------------------------------------------------- transform
def preprocess(image_u8: np.ndarray, size: int = 1024) -> torch.Tensor:
    """SAM2Transforms.__call__ (utils/transforms.py:15-40): ToTensor, Resize((size,size)) [bilinear, antialias],
    Normalize(ImageNet mean/std).  HxWx3 uint8 -> [1,3,size,size] f32."""
    x = torch.from_numpy(np.ascontiguousarray(image_u8)).permute(2, 0, 1).float().div(255.0)[None]
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=True)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (x - mean) / std


# ----------------------------------------------------------------------------------------------- Hiera trunk
def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _windows(x, ws):
    """backbones/utils.py:16-40 (grids here are always divisible by the window: no padding)."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def _unwindows(w, ws, H, W):
    """backbones/utils.py:43-63."""
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def _maxpool2(x):
    """do_pool with MaxPool2d(2,2) on a [B,H,W,C] tensor (hieradet.py:25-37)."""
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


def _sdpa(q, k, v):
    """F.scaled_dot_product_attention restated: softmax(q k^T / sqrt(d)) v.  [B,h,n,d]."""
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    return torch.softmax(s, dim=-1) @ v


def hiera_pos_embed(sd, grid: int) -> torch.Tensor:
    """Hiera._get_pos_embed (hieradet.py:264-272): bicubic background + tiled window embedding -> [1,g,g,C]."""
    pe = F.interpolate(sd["image_encoder.trunk.pos_embed"], size=(grid, grid), mode="bicubic")
    win = sd["image_encoder.trunk.pos_embed_window"]
    pe = pe + win.tile([x // y for x, y in zip(pe.shape, win.shape)])
    return pe.permute(0, 2, 3, 1)


def hiera_block(x, sd, p, spec, eps):
    """MultiScaleBlock.forward + MultiScaleAttention.forward (hieradet.py:134-166, 56-81)."""
    shortcut = x
    xn = _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    if spec.dim != spec.dim_out:
        shortcut = _lin(xn, sd, p + "proj")
        if spec.q_pool:
            shortcut = _maxpool2(shortcut)
    H = W = spec.grid_in
    ws = spec.window
    xw = _windows(xn, ws) if ws > 0 else xn
    B, h, w, _ = xw.shape
    qkv = _lin(xw, sd, p + "attn.qkv").reshape(B, h * w, 3, spec.heads, -1)
    q, k, v = qkv.unbind(2)
    if spec.q_pool:
        q = _maxpool2(q.reshape(B, h, w, -1))
        h, w = q.shape[1:3]
        q = q.reshape(B, h * w, spec.heads, -1)
    o = _sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(B, h, w, -1)
    o = _lin(o, sd, p + "attn.proj")
    if ws > 0:
        ws_out = ws // 2 if spec.q_pool else ws
        o = _unwindows(o, ws_out, spec.grid_out, spec.grid_out)
    x = shortcut + o
    xn2 = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    return x + _lin(F.gelu(_lin(xn2, sd, p + "mlp.layers.0")), sd, p + "mlp.layers.1")


def hiera(x, sd, cfg, taps=None):
    """Hiera.forward (hieradet.py:274-291): [1,3,S,S] -> list of stage outputs [1,g,g,C] (high to low resolution)."""
    t = "image_encoder.trunk."
    x = F.conv2d(x, sd[t + "patch_embed.proj.weight"], sd[t + "patch_embed.proj.bias"], stride=4, padding=3)
    x = x.permute(0, 2, 3, 1)
    x = x + hiera_pos_embed(sd, x.shape[1])
    if taps is not None:
        taps["patch"] = x.clone()
    outs = []
    ends = cfg.stage_ends()
    for i, spec in enumerate(cfg.blocks()):
        x = hiera_block(x, sd, f"{t}blocks.{i}.", spec, cfg.trunk_ln_eps)
        if taps is not None:
            taps[f"block{i}"] = x.clone()
        if i in ends:
            outs.append(x)
    return outs


def forward_image(pixels, sd, cfg, taps=None):
    """ImageEncoder + FpnNeck (backbones/image_encoder.py:29-134, nearest top-down on levels 2,3, scalp 1) +
    SAM2Base.forward_image conv_s0/conv_s1 (modeling/sam2_base.py:467-479) + the no_mem_embed add of
    SAM2ImagePredictor.set_image (sam2_image_predictor.py:118-127).
    -> image_embed [1,256,64,64], feat_s0 [1,32,256,256], feat_s1 [1,64,128,128]."""
    xs = [o.permute(0, 3, 1, 2) for o in hiera(pixels, sd, cfg, taps)]
    n = len(xs) - 1
    out = [None] * len(xs)
    prev = None
    for i in range(n, -1, -1):
        lat = F.conv2d(xs[i], sd[f"image_encoder.neck.convs.{n - i}.conv.weight"], sd[f"image_encoder.neck.convs.{n - i}.conv.bias"])
        if i in (2, 3) and prev is not None:
            prev = lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
        else:
            prev = lat
        out[i] = prev
    md = "sam_mask_decoder."
    feat_s0 = F.conv2d(out[0], sd[md + "conv_s0.weight"], sd[md + "conv_s0.bias"])
    feat_s1 = F.conv2d(out[1], sd[md + "conv_s1.weight"], sd[md + "conv_s1.bias"])
    image_embed = out[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)
    return image_embed, feat_s0, feat_s1


# ----------------------------------------------------------------------------------------------- prompt encoder
def _pe_encoding(coords01, gauss):
    """PositionEmbeddingRandom._pe_encoding (modeling/position_encoding.py:129-136)."""
    c = (2 * coords01 - 1) @ gauss
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd, size: int) -> torch.Tensor:
    """PromptEncoder.get_dense_pe (modeling/sam/prompt_encoder.py:69-79) -> [1,C,size,size]."""
    gauss = sd["sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones(size, size)
    y = (grid.cumsum(0) - 0.5) / size
    x = (grid.cumsum(1) - 0.5) / size
    return _pe_encoding(torch.stack([x, y], dim=-1), gauss).permute(2, 0, 1)[None]


def embed_points(points_xy, sd, image_size: int) -> torch.Tensor:
    """PromptEncoder._embed_points with one foreground point per prompt + the padding point
    (prompt_encoder.py:81-104): [B,2] pixel coords in the model frame -> sparse embeddings [B,2,C]."""
    pe = "sam_prompt_encoder."
    B = points_xy.shape[0]
    pts = torch.cat([points_xy[:, None, :] + 0.5, torch.zeros(B, 1, 2) + 0.5], dim=1)
    emb = _pe_encoding(pts / image_size, sd[pe + "pe_layer.positional_encoding_gaussian_matrix"])
    emb[:, 1, :] = 0.0
    emb[:, 1, :] += sd[pe + "not_a_point_embed.weight"][0]
    emb[:, 0, :] += sd[pe + "point_embeddings.1.weight"][0]
    return emb


# ----------------------------------------------------------------------------------------------- mask decoder
def _attn(q, k, v, sd, p, heads):
    """sam/transformer.py Attention.forward (:255-286)."""
    q, k, v = _lin(q, sd, p + ".q_proj"), _lin(k, sd, p + ".k_proj"), _lin(v, sd, p + ".v_proj")

    def split(x):
        b, n, c = x.shape
        return x.reshape(b, n, heads, c // heads).transpose(1, 2)

    o = _sdpa(split(q), split(k), split(v)).transpose(1, 2)
    return _lin(o.reshape(o.shape[0], o.shape[1], -1), sd, p + ".out_proj")


def _ln5(x, sd, name):
    return _ln(x, sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _mlp(x, sd, p, n, act=F.relu):
    for i in range(n):
        x = _lin(x, sd, f"{p}.layers.{i}")
        if i < n - 1:
            x = act(x)
    return x


def two_way_transformer(src, pos, tokens, sd, cfg):
    """TwoWayTransformer.forward / TwoWayAttentionBlock.forward (sam/transformer.py:90-134, 181-212).
    src, pos [B,HW,C]; tokens [B,T,C]."""
    p0 = "sam_mask_decoder.transformer."
    H = cfg.decoder_heads
    queries, keys = tokens, src
    for l in range(cfg.decoder_depth):
        p = f"{p0}layers.{l}."
        if l == 0:
            queries = _attn(queries, queries, queries, sd, p + "self_attn", H)
        else:
            q = queries + tokens
            queries = queries + _attn(q, q, queries, sd, p + "self_attn", H)
        queries = _ln5(queries, sd, p + "norm1")
        queries = queries + _attn(queries + tokens, keys + pos, keys, sd, p + "cross_attn_token_to_image", H)
        queries = _ln5(queries, sd, p + "norm2")
        queries = _ln5(queries + _mlp(queries, sd, p + "mlp", 2), sd, p + "norm3")
        keys = keys + _attn(keys + pos, queries + tokens, queries, sd, p + "cross_attn_image_to_token", H)
        keys = _ln5(keys, sd, p + "norm4")
    queries = queries + _attn(queries + tokens, keys + pos, keys, sd, p0 + "final_attn_token_to_image", H)
    return _ln5(queries, sd, p0 + "norm_final_attn"), keys


def _ln2d(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[None, :, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[None, :, None, None]


def predict(points_xy, image_embed, feat_s0, feat_s1, sd, cfg, taps=None):
    """SAM2ImagePredictor._predict with multimask_output=True (sam2_image_predictor.py:337-432) =
    prompt encoder + MaskDecoder.forward/predict_masks (sam/mask_decoder.py:110-245).
    points_xy [B,2] in model-frame pixels -> (low_res_masks [B,3,4g,4g] f32, iou [B,3])."""
    md = "sam_mask_decoder."
    B = points_xy.shape[0]
    g = image_embed.shape[-1]
    sparse = embed_points(points_xy, sd, cfg.image_size)
    out_tokens = torch.cat([sd[md + "obj_score_token.weight"], sd[md + "iou_token.weight"], sd[md + "mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tokens[None].expand(B, -1, -1), sparse], dim=1)
    src = image_embed + sd["sam_prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1)
    src = src.flatten(2).permute(0, 2, 1).expand(B, -1, -1)
    pos = dense_pe(sd, g).flatten(2).permute(0, 2, 1).expand(B, -1, -1)
    hs, keys = two_way_transformer(src, pos, tokens, sd, cfg)
    if taps is not None:
        taps["hs"], taps["keys"] = hs.clone(), keys.clone()
    iou_tok = hs[:, 1]
    mask_toks = hs[:, 2:2 + cfg.num_mask_tokens]
    src2 = keys.transpose(1, 2).reshape(B, -1, g, g)
    up = F.conv_transpose2d(src2, sd[md + "output_upscaling.0.weight"], sd[md + "output_upscaling.0.bias"], stride=2) + feat_s1
    up = F.gelu(_ln2d(up, sd[md + "output_upscaling.1.weight"], sd[md + "output_upscaling.1.bias"]))
    up = F.gelu(F.conv_transpose2d(up, sd[md + "output_upscaling.3.weight"], sd[md + "output_upscaling.3.bias"], stride=2) + feat_s0)
    hyper = torch.stack([_mlp(mask_toks[:, i], sd, f"{md}output_hypernetworks_mlps.{i}", 3) for i in range(cfg.num_mask_tokens)], 1)
    b, c, h, w = up.shape
    masks = (hyper @ up.reshape(b, c, h * w)).reshape(b, -1, h, w)
    iou = torch.sigmoid(_mlp(iou_tok, sd, md + "iou_prediction_head", 3))
    return masks[:, 1:], iou[:, 1:]


# ----------------------------------------------------------------------------------------------- AMG post-processing
def point_grid(n: int) -> np.ndarray:
    """utils/amg.py:181-189 build_point_grid."""
    off = 1 / (2 * n)
    side = np.linspace(off, 1 - off, n)
    return np.stack([np.tile(side[None, :], (n, 1)), np.tile(side[:, None], (1, n))], axis=-1).reshape(-1, 2)


def amg_points(n_per_side: int, H: int, W: int, image_size: int) -> np.ndarray:
    """The prompt coordinates exactly as the reference computes them: grid * (W,H) in f64 -> f32 tensor
    (automatic_mask_generator.py:264-265,308-310) -> /W, /H, * image_size in f32 (utils/transforms.py:59-65)."""
    pts = torch.as_tensor(point_grid(n_per_side) * np.array([[W, H]]), dtype=torch.float32)
    pts[..., 0] = pts[..., 0] / W
    pts[..., 1] = pts[..., 1] / H
    return (pts * image_size).numpy()


def box_nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """torchvision.ops.nms (CPU kernel) restated: stable descending sort, greedy, IoU = inter/(a_i+a_j-inter) > thr
    suppresses; returns kept indices in descending-score order."""
    order = np.argsort(-scores, kind="stable")
    b = boxes.astype(np.float32)
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)
    dead = np.zeros(len(b), bool)
    keep = []
    for ii, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        for j in order[ii + 1:]:
            if dead[j]:
                continue
            w = max(np.float32(0), min(b[i, 2], b[j, 2]) - max(b[i, 0], b[j, 0]))
            h = max(np.float32(0), min(b[i, 3], b[j, 3]) - max(b[i, 1], b[j, 1]))
            inter = np.float32(w * h)
            if inter / np.float32(area[i] + area[j] - inter) > np.float32(thr):
                dead[j] = True
    return np.array(keep, np.int64)


def mask_boxes(masks: np.ndarray) -> np.ndarray:
    """utils/amg.py:305-348 batched_mask_to_box: XYXY (inclusive max), [0,0,0,0] for an empty mask."""
    out = np.zeros((masks.shape[0], 4), np.int64)
    for i, m in enumerate(masks):
        ys, xs = np.nonzero(m)
        if len(ys):
            out[i] = (xs.min(), ys.min(), xs.max(), ys.max())
    return out


def upsample_bilinear(low: np.ndarray, H: int, W: int) -> np.ndarray:
    """F.interpolate(mode="bilinear", align_corners=False) (utils/transforms.py:117) restated with a fixed f32 operation
    order (ATen upsample_bilinear2d: src = scale*(dst+0.5)-0.5 clamped at 0, i0 = int(src), lambda = src - i0,
    out = (1-ly)*((1-lx)*a + lx*b) + ly*((1-lx)*c + lx*d)); ATen's own CPU kernel may contract some of these into FMAs, so the
    two can differ in the last bit — the thresholded masks are pinned to the reference within a few pixels instead."""
    f = np.float32
    h, w = low.shape[-2:]

    def axis(n_out, n_in):
        scale = f(n_in) / f(n_out)
        src = (scale * (np.arange(n_out, dtype=f) + f(0.5))).astype(f) - f(0.5)
        src = np.maximum(src, f(0)).astype(f)
        i0 = np.minimum(src.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        lam = (src - i0.astype(f)).astype(f)
        return i0, i1, lam, (f(1) - lam).astype(f)

    y0, y1, ly, hy = axis(H, h)
    x0, x1, lx, hx = axis(W, w)
    low = low.astype(f, copy=False)
    out = np.empty(low.shape[:-2] + (H, W), f)
    flat, oflat = low.reshape(-1, h, w), out.reshape(-1, H, W)
    for i in range(flat.shape[0]):
        m = flat[i]
        top = (hx * m[y0][:, x0]).astype(f) + (lx * m[y0][:, x1]).astype(f)
        bot = (hx * m[y1][:, x0]).astype(f) + (lx * m[y1][:, x1]).astype(f)
        oflat[i] = (hy[:, None] * top).astype(f) + (ly[:, None] * bot).astype(f)
    return out


def amg_postprocess(low_res, iou, H, W, pred_iou_thresh=0.8, stability_thresh=0.95, offset=1.0, box_nms_thresh=0.7):
    """SAM2AutomaticMaskGenerator._process_batch/_process_crop for a single full-image crop
    (automatic_mask_generator.py:251-292, 294-375) applied to ALL prompts at once (the per-64-prompt batching of the
    reference only bounds memory: every filter is per mask).  low_res [P,3,h,w] logits, iou [P,3].
    -> dict(masks bool [K,H,W], iou [K], stability [K], boxes [K,4], src [K] = index into the flattened P*3 list),
    in the order the reference returns them (descending predicted IoU after box NMS)."""
    P = low_res.shape[0]
    iou = iou.flatten().float()
    src = torch.arange(P * 3)
    keep = iou > pred_iou_thresh
    iou, src = iou[keep], src[keep]
    # (the reference up-samples every mask before the IoU filter; per-mask results are the same)
    masks = torch.from_numpy(upsample_bilinear(low_res.float().flatten(0, 1)[keep].numpy(), H, W))
    inter = (masks > offset).flatten(1).sum(1).to(torch.int32)
    union = (masks > -offset).flatten(1).sum(1).to(torch.int32)
    stab = inter / union
    keep = stab >= stability_thresh
    masks, iou, src, stab = masks[keep], iou[keep], src[keep], stab[keep]
    bin_ = (masks > 0.0).numpy()
    boxes = mask_boxes(bin_)
    # is_box_near_crop_edge (amg.py:79-90): the crop is the whole image, so no box is dropped
    order = box_nms(boxes.astype(np.float32), iou.numpy(), box_nms_thresh)
    return dict(masks=bin_[order], iou=iou.numpy()[order], stability=stab.numpy()[order], boxes=boxes[order],
                src=src.numpy()[order])


def generate(image_u8, sd, cfg, points_per_side=16, pred_iou_thresh=0.8, stability_thresh=0.95, offset=1.0, box_nms_thresh=0.7):
    """SAM2AutomaticMaskGenerator.generate (automatic_mask_generator.py:170-222), crop_n_layers 0."""
    H, W = image_u8.shape[:2]
    with torch.no_grad():
        emb, s0, s1 = forward_image(preprocess(image_u8, cfg.image_size), sd, cfg)
        pts = torch.from_numpy(amg_points(points_per_side, H, W, cfg.image_size))
        lows, ious = [], []
        for b in range(0, pts.shape[0], 64):
            lo, io = predict(pts[b:b + 64], emb, s0, s1, sd, cfg)
            lows.append(lo); ious.append(io)
    return amg_postprocess(torch.cat(lows), torch.cat(ious), H, W, pred_iou_thresh, stability_thresh, offset, box_nms_thresh)
