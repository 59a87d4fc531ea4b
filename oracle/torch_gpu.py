"""TEST / BENCH INFRASTRUCTURE — the reference's DEPLOYMENT path restated in plain PyTorch for the GPU: what OVO runs on a CUDA
device (bf16 autocast, ovo/entities/ovomapping.py:166; SDPA pe.py:145-147; F.linear pe.py:125,150; torch ops of
ovo/utils/geometry_utils.py and the per-mask Python loop of ovo/entities/ovo.py:240-282; torch.mm of clip_utils.py:16-19).
Only bench.py's `gpu_baseline` leg runs it (the GPU-vs-GPU bar of SURVEY 8d): never the product, never the thing shipped.
/root/reference does not exist on the GPU box, so nothing here imports it."""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ encoder (pe.py)
def rope_table(grid: int, head_dim: int, device):
    half = head_dim // 2
    nf = half // 2
    theta = 1.0 / (10000 ** (torch.arange(0, half, 2, device=device)[:nf].float() / half))
    ys, xs = torch.meshgrid(torch.arange(grid, device=device), torch.arange(grid, device=device), indexing="ij")
    ang = torch.cat([(xs.reshape(-1, 1).float() + 1) * theta[None], (ys.reshape(-1, 1).float() + 1) * theta[None]], dim=1)
    ang = torch.cat([torch.zeros(1, ang.shape[1], device=device), ang], dim=0)
    return ang.cos(), ang.sin()


def apply_rope(t, cos, sin):
    a, b = t[..., 0::2], t[..., 1::2]
    return torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)


def resblock(x, W, pfx, heads, eps, rope):
    """ResidualAttentionBlock (pe.py:216-225) with SelfAttention (pe.py:123-150): F.linear + RoPE + SDPA + F.linear."""
    B, S, D = x.shape
    hd = D // heads
    h = F.layer_norm(x, (D,), W[pfx + "ln_1.weight"], W[pfx + "ln_1.bias"], eps)
    qkv = F.linear(h, W[pfx + "attn.in_proj_weight"], W[pfx + "attn.in_proj_bias"])
    q, k, v = (t.view(B, S, heads, hd).transpose(1, 2) for t in qkv.split(D, dim=-1))
    q, k = apply_rope(q, *rope).to(v.dtype), apply_rope(k, *rope).to(v.dtype)
    a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, D)
    x = x + F.linear(a, W[pfx + "attn.out_proj.weight"], W[pfx + "attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), W[pfx + "ln_2.weight"], W[pfx + "ln_2.bias"], eps)
    h = F.gelu(F.linear(h, W[pfx + "mlp.c_fc.weight"], W[pfx + "mlp.c_fc.bias"]))
    return x + F.linear(h, W[pfx + "mlp.c_proj.weight"], W[pfx + "mlp.c_proj.bias"])


def vit_forward_features(pixels, W, cfg):
    """VisionTransformer.forward_features(norm=True) (pe.py:499-533) — call under torch.autocast('cuda', torch.bfloat16)."""
    n = pixels.shape[0]
    x = F.conv2d(pixels, W["visual.conv1.weight"], stride=cfg.patch_size)
    x = x.permute(0, 2, 3, 1).reshape(n, -1, cfg.width)
    cls = W["visual.class_embedding"].view(1, 1, -1).expand(n, -1, -1)
    x = torch.cat([cls.to(x.dtype), x], dim=1) + W["visual.positional_embedding"][None]
    x = F.layer_norm(x, (cfg.width,), W["visual.ln_pre.weight"], W["visual.ln_pre.bias"], cfg.ln_eps)
    rope = rope_table(cfg.image_size // cfg.patch_size, cfg.width // cfg.heads, pixels.device)
    for i in range(cfg.layers):
        x = resblock(x, W, f"visual.transformer.resblocks.{i}.", cfg.heads, cfg.ln_eps, rope)
    return F.layer_norm(x, (cfg.width,), W["visual.ln_post.weight"], W["visual.ln_post.bias"], cfg.ln_eps)


# ------------------------------------------------------------------------------------------------ association (geometry_utils.py, ovo.py)
def associate(xyz, ins_ids, depth, seg_map, c2w, K, match_th=0.05, track_th=100, next_ins_id=0):
    """OVO._match_and_track_instances + _track_objects (ovo.py:204-229, 240-282) with torch ops on the device, structured like
    the reference: frustum cull (geometry_utils.py:99-129,252-277), depth filter (:92-96), projection / depth match (:26-89),
    seg lookup, then the per-mask Python loop with its .item() synchronisations.  Returns (ins_ids_new, n_matched, next_ins_id)."""
    dev = xyz.device
    h, w = depth.shape
    valid = depth[depth > 0]
    dmin, dmax = valid.min(), valid.max()
    px = torch.tensor([0, w, 0, w, 0, w, 0, w], device=dev, dtype=torch.float32)
    py = torch.tensor([0, 0, h, h, 0, 0, h, h], device=dev, dtype=torch.float32)
    pz = torch.cat([dmin.expand(4), dmax.expand(4)])
    cam = torch.stack([(px - K[0, 2]) * pz / K[0, 0], (py - K[1, 2]) * pz / K[1, 1], pz, torch.ones_like(pz)], dim=1)
    c = (cam @ c2w.T)[:, :3]
    lo, hi = c.min(0).values, c.max(0).values
    m = ((xyz >= lo) & (xyz <= hi)).all(dim=1)
    pairs = [(2, 0, 1, 0), (6, 4, 5, 4), (4, 0, 2, 0), (7, 3, 1, 3), (5, 1, 3, 1), (6, 2, 0, 2)]
    for i, (a, b, cc, d) in enumerate(pairs):
        nrm = torch.linalg.cross(c[a] - c[b], c[cc] - c[d])
        m &= (xyz @ nrm - (nrm * c[i]).sum()) <= 0
    idx = torch.nonzero(m).squeeze(1)
    # depth filter (7x7 gaussian high-pass, geometry_utils.py:92-96)
    k1 = torch.exp(-0.5 * (torch.linspace(-3, 3, 7, device=dev) / 2.5) ** 2)
    k1 = k1 / k1.sum()
    blur = F.conv2d(F.pad(depth[None, None], (3, 3, 3, 3), mode="reflect"), (k1[:, None] * k1[None, :])[None, None])[0, 0]
    d = torch.where((depth - blur).abs() > 0.05, torch.full_like(depth, -1.0), depth)
    pts = xyz[idx]
    w2c = torch.linalg.inv(c2w)
    loc = torch.cat([pts, torch.ones(pts.shape[0], 1, device=dev)], dim=1) @ w2c.T
    loc = loc[:, :3] / loc[:, 3:4]
    ph = loc @ K.T
    uv = torch.round(ph[:, :2] / ph[:, 2:3]).long()
    inb = (uv[:, 0] >= 0) & (uv[:, 0] < w) & (uv[:, 1] >= 0) & (uv[:, 1] < h)
    uvc = uv.clamp_min(0)
    uvc[:, 0].clamp_max_(w - 1); uvc[:, 1].clamp_max_(h - 1)
    dz = d[uvc[:, 1], uvc[:, 0]]
    ok = inb & ((loc[:, 2] - dz).abs() < match_th) & (dz != 0)
    matched = idx[ok]
    seg_of = seg_map[uvc[ok][:, 1], uvc[ok][:, 0]]
    out = ins_ids.clone()
    for map_idx in range(int(seg_map.max().item()) + 1):          # ovo.py:255-280
        map_points = matched[seg_of == map_idx]
        if len(map_points) > track_th:
            (seg_map == map_idx).sum().item()
            assigned = out[map_points] > -1
            map_ins = -1
            if assigned.sum().item() > track_th:
                map_ins = torch.mode(out[map_points[assigned]]).values.item()
            elif (~assigned).sum().item() > track_th:
                map_ins = next_ins_id
                next_ins_id += 1
            if map_ins > -1:
                out[map_points[~assigned]] = map_ins
    return out, int(matched.shape[0]), next_ins_id
