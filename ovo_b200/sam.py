"""Host side of the SAM-2 mask proposal: packs a SAM-2 state_dict (the reference's own key names) into
device buffers and drives libovo_b200 through the C ABI.

Mirrors `SAM2ImagePredictor.set_image/_predict` (thirdParty/segment-anything-2/sam2/sam2_image_predictor.py:86-127,337-432),
`SAM2AutomaticMaskGenerator.generate` (sam2/automatic_mask_generator.py:170-222) as OVO wires it
(ovo/utils/segment_utils.py:269-309) and `MaskGenerator.segment` (ovo/entities/mask_generator.py:102-120)."""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import AmgParams, HieraBlock, SamAttn, SamCfg, SamDecLayer, SamWeights, check, ptr, stream_ptr
from .sam_config import SamConfig


class Sam2:
    """Device-resident SAM-2 image encoder + prompt/mask decoder + automatic mask generator."""

    def __init__(self, cfg: SamConfig, state_dict: dict, max_h: int = 480, max_w: int = 640, max_prompts: int = 256, device="cuda",
                 max_batch: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("ovo_b200.Sam2 needs a CUDA device (no CPU fallback)")
        self.cfg, self.device = cfg, torch.device(device)
        self.lib = _lib.lib()
        self._keep = []
        self.g = cfg.image_size // 16
        w = self._pack(state_dict)
        c = SamCfg()
        c.image_size, c.n_blocks, c.embed_dim = cfg.image_size, len(cfg.blocks()), cfg.embed_dim
        for i, e in enumerate(cfg.stage_ends()):
            c.stage_end[i] = e
        c.decoder_depth, c.trunk_ln_eps, c.max_batch = cfg.decoder_depth, cfg.trunk_ln_eps, int(max_batch)
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ovo_sam_create(C.byref(c), C.byref(w), max_h, max_w, max_prompts, C.byref(h)), "ovo_sam_create")
        self.handle = h
        self.max_prompts = max_prompts

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ovo_sam_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _dev(self, t, dtype):
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(t)
        return ptr(t)

    def _lin(self, sd, name):
        return self._dev(sd[name + ".weight"], torch.bfloat16), self._dev(sd[name + ".bias"], torch.float32)

    def _attn(self, sd, name) -> SamAttn:
        a = SamAttn()
        a.q_w, a.q_b = self._lin(sd, name + ".q_proj"); a.k_w, a.k_b = self._lin(sd, name + ".k_proj")
        a.v_w, a.v_b = self._lin(sd, name + ".v_proj"); a.o_w, a.o_b = self._lin(sd, name + ".out_proj")
        return a

    def _pack(self, sd) -> SamWeights:
        cfg = self.cfg
        f32, bf = torch.float32, torch.bfloat16
        sd = {k: v.detach().float().cpu() for k, v in sd.items()}
        w = SamWeights()
        t = "image_encoder.trunk."
        E = cfg.embed_dim
        kpad = 152                                           # 3*7*7 = 147 rounded up to a multiple of 8 (16-byte TMA rows)
        pw = torch.zeros(E, kpad)
        pw[:, :147] = sd[t + "patch_embed.proj.weight"].reshape(E, 147)
        w.patch_w, w.patch_kpad, w.patch_b = self._dev(pw, bf), kpad, self._dev(sd[t + "patch_embed.proj.bias"], f32)
        # Hiera._get_pos_embed (hieradet.py:264-272): a function of the weights only -> evaluated once at load time
        g0 = cfg.image_size // 4
        pe = F.interpolate(sd[t + "pos_embed"], size=(g0, g0), mode="bicubic")
        win = sd[t + "pos_embed_window"]
        pe = pe + win.tile([x // y for x, y in zip(pe.shape, win.shape)])
        w.pos = self._dev(pe.permute(0, 2, 3, 1).reshape(g0 * g0, E), f32)
        specs = cfg.blocks()
        arr = (HieraBlock * len(specs))()
        for i, b in enumerate(specs):
            p = f"{t}blocks.{i}."
            hb = arr[i]
            hb.dim, hb.dim_out, hb.heads, hb.window, hb.q_pool, hb.grid_in = b.dim, b.dim_out, b.heads, b.window, int(b.q_pool), b.grid_in
            hb.norm1_w, hb.norm1_b = self._dev(sd[p + "norm1.weight"], f32), self._dev(sd[p + "norm1.bias"], f32)
            hb.qkv_w, hb.qkv_b = self._lin(sd, p + "attn.qkv")
            hb.proj_w, hb.proj_b = self._lin(sd, p + "attn.proj")
            hb.norm2_w, hb.norm2_b = self._dev(sd[p + "norm2.weight"], f32), self._dev(sd[p + "norm2.bias"], f32)
            hb.fc1_w, hb.fc1_b = self._lin(sd, p + "mlp.layers.0")
            hb.fc2_w, hb.fc2_b = self._lin(sd, p + "mlp.layers.1")
            if b.dim != b.dim_out:
                hb.short_w, hb.short_b = self._lin(sd, p + "proj")
        self._keep.append(arr)
        w.blocks = arr
        # neck: convs[j] belongs to stage 3-j (image_encoder.py:108-113); conv_s0/conv_s1 folded in f64
        md = "sam_mask_decoder."
        nw = [sd[f"image_encoder.neck.convs.{j}.conv.weight"].double().flatten(1) for j in range(4)]
        nb = [sd[f"image_encoder.neck.convs.{j}.conv.bias"].double() for j in range(4)]
        w.neck3_w, w.neck3_b = self._dev(nw[0].float(), bf), self._dev(nb[0].float(), f32)
        w.neck2_w = self._dev(nw[1].float(), bf)
        w.neck2_b = self._dev((nb[1] + sd["no_mem_embed"].double().reshape(-1)).float(), f32)   # sam2_image_predictor.py:118-121
        ws1, bs1 = sd[md + "conv_s1.weight"].double().flatten(1), sd[md + "conv_s1.bias"].double()
        ws0, bs0 = sd[md + "conv_s0.weight"].double().flatten(1), sd[md + "conv_s0.bias"].double()
        w.s1_w, w.s1_b = self._dev((ws1 @ nw[2]).float(), bf), self._dev((ws1 @ nb[2] + bs1).float(), f32)
        w.s0_w, w.s0_b = self._dev((ws0 @ nw[3]).float(), bf), self._dev((ws0 @ nb[3] + bs0).float(), f32)
        pe_ = "sam_prompt_encoder."
        gauss = sd[pe_ + "pe_layer.positional_encoding_gaussian_matrix"]
        w.gauss = self._dev(gauss, f32)
        w.point_embed = self._dev(sd[pe_ + "point_embeddings.1.weight"][0], f32)
        w.not_a_point = self._dev(sd[pe_ + "not_a_point_embed.weight"][0], f32)
        w.no_mask_embed = self._dev(sd[pe_ + "no_mask_embed.weight"][0], f32)
        # PromptEncoder.get_dense_pe (prompt_encoder.py:69-79, position_encoding.py:129-149): weights only
        g = self.g
        grid = torch.ones(g, g)
        ye, xe = (grid.cumsum(0) - 0.5) / g, (grid.cumsum(1) - 0.5) / g
        cc = (2 * torch.stack([xe, ye], dim=-1) - 1) @ gauss
        cc = 2 * math.pi * cc
        w.dense_pe = self._dev(torch.cat([torch.sin(cc), torch.cos(cc)], dim=-1).reshape(g * g, -1), f32)
        w.out_tokens = self._dev(torch.cat([sd[md + "obj_score_token.weight"], sd[md + "iou_token.weight"], sd[md + "mask_tokens.weight"]], 0), f32)
        layers = (SamDecLayer * cfg.decoder_depth)()
        for l in range(cfg.decoder_depth):
            p = f"{md}transformer.layers.{l}."
            L = layers[l]
            L.self_attn = self._attn(sd, p + "self_attn")
            L.t2i = self._attn(sd, p + "cross_attn_token_to_image")
            L.i2t = self._attn(sd, p + "cross_attn_image_to_token")
            for j in range(4):
                L.norm_w[j] = self._dev(sd[f"{p}norm{j + 1}.weight"], f32)
                L.norm_b[j] = self._dev(sd[f"{p}norm{j + 1}.bias"], f32)
            L.mlp0_w, L.mlp0_b = self._lin(sd, p + "mlp.layers.0")
            L.mlp1_w, L.mlp1_b = self._lin(sd, p + "mlp.layers.1")
        self._keep.append(layers)
        w.layers = layers
        w.final_attn = self._attn(sd, md + "transformer.final_attn_token_to_image")
        w.norm_final_w = self._dev(sd[md + "transformer.norm_final_attn.weight"], f32)
        w.norm_final_b = self._dev(sd[md + "transformer.norm_final_attn.bias"], f32)
        # ConvTranspose2d k2 s2 (weight [in, out, ky, kx]) as a GEMM with rows (ky*2+kx)*out + o
        u0, u1 = sd[md + "output_upscaling.0.weight"], sd[md + "output_upscaling.3.weight"]
        w.up0_w = self._dev(u0.permute(2, 3, 1, 0).reshape(4 * u0.shape[1], u0.shape[0]), bf)
        w.up0_b = self._dev(sd[md + "output_upscaling.0.bias"].repeat(4), f32)
        w.up_ln_w = self._dev(sd[md + "output_upscaling.1.weight"], f32)
        w.up_ln_b = self._dev(sd[md + "output_upscaling.1.bias"], f32)
        w.up1_w = self._dev(u1.permute(2, 3, 1, 0).reshape(4 * u1.shape[1], u1.shape[0]), bf)
        w.up1_b = self._dev(sd[md + "output_upscaling.3.bias"].repeat(4), f32)
        for i in range(4):
            for j in range(3):
                w.hyper_w[i][j], w.hyper_b[i][j] = self._lin(sd, f"{md}output_hypernetworks_mlps.{i}.layers.{j}")
        for j in range(3):
            w.iou_w[j], w.iou_b[j] = self._lin(sd, f"{md}iou_prediction_head.layers.{j}")
        return w

    # ------------------------------------------------------------------ calls
    def _taps(self, want):
        g, dev = self.g, self.device
        if not want:
            return None, None, None
        return (torch.empty(g * g, 256, device=dev), torch.empty(16 * g * g, 32, device=dev), torch.empty(4 * g * g, 64, device=dev))

    def set_image(self, rgb_u8: torch.Tensor, taps: bool = False, n_blocks: int = -1):
        """SAM2ImagePredictor.set_image.  rgb uint8 [H,W,3].  taps=True returns (pixels [3,S,S], image_embed [g*g,256],
        feat_s0 [16g*g,32], feat_s1 [4g*g,64]) — token-major views of the reference's NCHW features."""
        rgb = rgb_u8.to(self.device, torch.uint8).contiguous()
        H, W, _ = rgb.shape
        S = self.cfg.image_size
        px = torch.empty(3, S, S, device=self.device) if taps else None
        emb, s0, s1 = self._taps(taps and n_blocks < 0)
        blk = None
        if n_blocks >= 0:
            spec = self.cfg.blocks()[n_blocks - 1] if n_blocks > 0 else None
            gsz, dim = (spec.grid_out, spec.dim_out) if spec else (S // 4, self.cfg.embed_dim)
            blk = torch.empty(gsz * gsz, dim, device=self.device)
        check(self.lib.ovo_sam_set_image(self.handle, ptr(rgb), H, W, ptr(px), ptr(emb), ptr(s0), ptr(s1), n_blocks, ptr(blk),
                                         stream_ptr(self.device)), "ovo_sam_set_image")
        self._rgb = rgb
        return (px, emb, s0, s1) if n_blocks < 0 else blk

    def set_pixels(self, pixels: torch.Tensor, n_blocks: int = -1):
        """Test tap: run the trunk (and neck) from normalised pixels [3,S,S]."""
        pixels = pixels.to(self.device, torch.float32).contiguous()
        S = self.cfg.image_size
        emb, s0, s1 = self._taps(n_blocks < 0)
        blk = None
        if n_blocks >= 0:
            spec = self.cfg.blocks()[n_blocks - 1] if n_blocks > 0 else None
            gsz, dim = (spec.grid_out, spec.dim_out) if spec else (S // 4, self.cfg.embed_dim)
            blk = torch.empty(gsz * gsz, dim, device=self.device)
        check(self.lib.ovo_sam_set_pixels(self.handle, ptr(pixels), ptr(emb), ptr(s0), ptr(s1), n_blocks, ptr(blk), stream_ptr(self.device)),
              "ovo_sam_set_pixels")
        return (emb, s0, s1) if n_blocks < 0 else blk

    def set_images(self, rgb_u8: torch.Tensor):
        """set_image for a batch [n,H,W,3] (n <= max_batch) in ONE trunk pass; `select_image(i)` chooses the frame the
        decoder (`predict`) works on."""
        rgb = rgb_u8.to(self.device, torch.uint8).contiguous()
        n, H, W, _ = rgb.shape
        check(self.lib.ovo_sam_set_images(self.handle, ptr(rgb), n, H, W, stream_ptr(self.device)), "ovo_sam_set_images")
        self._rgb = rgb

    def select_image(self, index: int):
        check(self.lib.ovo_sam_select_image(self.handle, int(index)), "ovo_sam_select_image")

    def predict(self, points: torch.Tensor):
        """SAM2ImagePredictor._predict (one foreground point per prompt, multimask_output=True, return_logits=True).
        points f32 [P,2] in model-frame pixels -> (low_res_masks [P,3,4g,4g] f32 — NOT clamped —, iou [P,3])."""
        pts = points.to(self.device, torch.float32).contiguous()
        P = pts.shape[0]
        low = torch.empty(P, 3, 4 * self.g, 4 * self.g, device=self.device)
        iou = torch.empty(P, 3, device=self.device)
        check(self.lib.ovo_sam_predict(self.handle, ptr(pts), P, ptr(low), ptr(iou), stream_ptr(self.device)), "ovo_sam_predict")
        return low, iou

    @staticmethod
    def amg_params(points_per_side=16, pred_iou_thresh=0.8, stability_score_thresh=0.95, stability_score_offset=1.0,
                   box_nms_thresh=0.7, nms_iou_th=0.8, nms_score_th=0.7, nms_inner_th=0.5) -> AmgParams:
        """Defaults = OVO's wiring (segment_utils.py:296-302, automatic_mask_generator.py:40-56, mask_generator.py:25-27)."""
        return AmgParams(points_per_side, pred_iou_thresh, stability_score_thresh, stability_score_offset, box_nms_thresh,
                         nms_iou_th, nms_score_th, nms_inner_th)

    def postprocess(self, low: torch.Tensor, iou: torch.Tensor, H: int, W: int, prm: AmgParams):
        """The AMG filters on given logits -> dict(masks uint8 [K,H,W], iou, stability, boxes XYXY, src)."""
        P, _, h, w = low.shape
        n = P * 3
        masks = torch.empty(n, H, W, device=self.device, dtype=torch.uint8)
        io = torch.empty(n, device=self.device); st = torch.empty(n, device=self.device)
        boxes = torch.empty(n, 4, device=self.device, dtype=torch.int32); src = torch.empty(n, device=self.device, dtype=torch.int32)
        k = C.c_int(0)
        check(self.lib.ovo_sam_postprocess(self.handle, ptr(low.contiguous()), ptr(iou.contiguous()), P, h, w, H, W, C.byref(prm),
                                           ptr(masks), ptr(io), ptr(st), ptr(boxes), ptr(src), n, C.byref(k), stream_ptr(self.device)),
              "ovo_sam_postprocess")
        K = k.value
        return dict(masks=masks[:K], iou=io[:K], stability=st[:K], boxes=boxes[:K], src=src[:K])

    def override_logits(self, low: torch.Tensor | None, iou: torch.Tensor | None = None) -> None:
        """Measurement / test aid (ovo_sam_override_logits): `generate` keeps running the network but post-processes these
        logits [P,3,4g,4g] / predicted IoUs [P,3] instead (None removes the override).  See `synthetic_logits`."""
        if low is None:
            self._override = None
            check(self.lib.ovo_sam_override_logits(self.handle, None, None, 0), "ovo_sam_override_logits")
            return
        low = low.to(self.device, torch.float32).contiguous()
        iou = iou.to(self.device, torch.float32).contiguous()
        assert low.dim() == 4 and low.shape[1] == 3 and low.shape[2] == 4 * self.g and iou.shape == (low.shape[0], 3)
        self._override = (low, iou)
        check(self.lib.ovo_sam_override_logits(self.handle, ptr(low), ptr(iou), low.shape[0]), "ovo_sam_override_logits")

    def synthetic_logits(self, points_per_side: int = 16, keep_frac: float = 0.55, seed: int = 0):
        """Plausible decoder outputs for a benchmark without checkpoints: every grid prompt proposes three nested elliptical
        masks around its point (logit +-8 with a soft edge, like a confident SAM-2), a seeded `keep_frac` of the prompts gets a
        high predicted IoU for ONE of its scales, the rest are rejected by the stock pred_iou threshold.  With the stock AMG /
        OVO thresholds (0.8 / 0.95 / box-NMS 0.7, mask NMS 0.8 / 0.7 / 0.5) 50-150 masks survive."""
        gen = torch.Generator().manual_seed(seed)
        n, S = points_per_side, 4 * self.g
        c = (torch.arange(n, dtype=torch.float32) + 0.5) / n * S
        cy, cx = torch.meshgrid(c, c, indexing="ij")
        yy, xx = torch.meshgrid(torch.arange(S, dtype=torch.float32), torch.arange(S, dtype=torch.float32), indexing="ij")
        P = n * n
        low = torch.empty(P, 3, S, S)
        rad = torch.tensor([0.45, 0.8, 1.35]) * (S / n)
        ecc = 0.6 + 0.8 * torch.rand(P, generator=gen)
        for k in range(3):
            d = torch.sqrt(((xx[None] - cx.reshape(-1, 1, 1)) * ecc[:, None, None]) ** 2 + ((yy[None] - cy.reshape(-1, 1, 1)) / ecc[:, None, None]) ** 2)
            low[:, k] = ((rad[k] - d) * 20.0).clamp(-8, 8)
        iou = torch.full((P, 3), 0.3)
        chosen = torch.rand(P, generator=gen) < keep_frac
        scale = torch.randint(0, 3, (P,), generator=gen)
        iou[torch.nonzero(chosen).squeeze(1), scale[chosen]] = 0.93
        return low, iou

    def generate(self, rgb_u8: torch.Tensor, prm: AmgParams = None, max_masks: int = 256):
        """MaskGenerator.segment: rgb uint8 [H,W,3] -> (seg_map int32 [H,W], binary_maps bool [M,H,W])."""
        prm = prm or self.amg_params()
        rgb = rgb_u8.to(self.device, torch.uint8).contiguous()
        H, W, _ = rgb.shape
        seg = torch.full((H, W), -1, device=self.device, dtype=torch.int32)
        maps = torch.empty(max_masks, H, W, device=self.device, dtype=torch.uint8)
        m = C.c_int(0)
        check(self.lib.ovo_sam_generate(self.handle, ptr(rgb), H, W, C.byref(prm), ptr(seg), ptr(maps), max_masks, C.byref(m),
                                        stream_ptr(self.device)), "ovo_sam_generate")
        return seg, maps[: m.value].bool()

    def generate_batch(self, rgb_u8: torch.Tensor, prm: AmgParams = None, max_masks: int = 256):
        """`generate` for n <= max_batch frames [n,H,W,3] with one batched trunk pass (MaskGenerator.precompute / replay).
        -> list of (seg_map int32 [H,W], binary_maps bool [M,H,W])."""
        prm = prm or self.amg_params()
        rgb = rgb_u8.to(self.device, torch.uint8).contiguous()
        n, H, W, _ = rgb.shape
        seg = torch.full((n, H, W), -1, device=self.device, dtype=torch.int32)
        maps = torch.empty(n, max_masks, H, W, device=self.device, dtype=torch.uint8)
        m = (C.c_int * n)()
        check(self.lib.ovo_sam_generate_batch(self.handle, ptr(rgb), n, H, W, C.byref(prm), ptr(seg), ptr(maps), max_masks, m,
                                              stream_ptr(self.device)), "ovo_sam_generate_batch")
        return [(seg[i], maps[i, : m[i]].bool()) for i in range(n)]
