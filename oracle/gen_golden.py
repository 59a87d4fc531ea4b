"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on seeded synthetic inputs.

    python -m oracle.gen_golden          (build container only; the GPU box has no /root/reference)

Inputs and weights are NOT stored: they are regenerated in the tests from seeds
(ovo_b200.synth + ovo_b200.encoder.random_state_dict), only the reference's outputs are committed.
The reference model is pe.CLIP built from a small PEConfig (same code path as PE-Core-L14-336: cls token,
abs pos-emb, 2D RoPE, attention pooler; head_dim 64) with our seeded weights loaded into it.
"""
import os
import sys
import tempfile

import numpy as np
import torch

from . import ref_harness as rh
from . import encoder as OE

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# the small encoder used by every fixture
TINY = dict(width=128, layers=2, heads=2, mlp_width=512, output_dim=64, text_width=128, text_heads=2, text_layers=2,
            text_mlp_width=512, text_output_dim=64, vocab_size=1024)
TEXT_TOKENS = np.array([[1000, 5, 17, 1001] + [0] * 28,
                        [1000, 900, 3, 44, 2, 1001] + [0] * 26,
                        [1000, 7, 1001] + [0] * 29], np.int64)
QUERIES_TOK = TEXT_TOKENS  # three "queries" expressed directly as token ids (tiny vocab)


def tiny_cfg():
    from ovo_b200.encoder import EncoderConfig
    return EncoderConfig(**TINY)


def build_reference_model(seed=0):
    """Reference pe.CLIP (tiny config) carrying OUR seeded weights."""
    from ovo_b200.encoder import random_state_dict
    cfg = tiny_cfg()
    rh.setup_paths()
    from core.vision_encoder.config import PEConfig, PETextConfig
    import core.vision_encoder.pe as pe
    vc = PEConfig(image_size=cfg.image_size, patch_size=cfg.patch_size, width=cfg.width, layers=cfg.layers,
                  heads=cfg.heads, mlp_ratio=cfg.mlp_width / cfg.width, pool_type="attn", output_dim=cfg.output_dim,
                  use_cls_token=True, attn_pooler_heads=cfg.heads)
    tc = PETextConfig(context_length=cfg.text_ctx, width=cfg.text_width, heads=cfg.text_heads, layers=cfg.text_layers,
                      output_dim=cfg.text_output_dim, mlp_ratio=cfg.text_mlp_width / cfg.text_width,
                      vocab_size=cfg.vocab_size)
    torch.manual_seed(123)
    model = pe.CLIP(vc, tc).eval()
    sd = random_state_dict(cfg, seed=seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # keys we do not set never influence the TextRegion / text outputs: pooler probe/ln/mlp/q,k
    assert all(k.startswith("visual.attn_pool.") or k == "logit_scale" for k in missing), missing
    return model, sd, cfg


def masks_for(h, w):
    from ovo_b200 import synth
    seg, bm = synth.grid_masks(h, w, rows=3, cols=4)
    bm = np.concatenate([bm, np.zeros((1, h, w), bool)])
    bm[-1, h // 2: h // 2 + 3, 5:9] = True          # too small to touch a token -> NaN row in the reference
    return bm


def gen_encoder():
    from ovo_b200 import synth
    model, sd, cfg = build_reference_model()
    tr = rh.build_textregion(model)
    out = {}
    for tag, (h, w) in {"a": (480, 640), "b": (968, 1296)}.items():
        img = synth.rgb(h, w, seed=3)
        bm = masks_for(h, w)
        imt = torch.from_numpy(img.transpose(2, 0, 1).copy()).float()
        with torch.no_grad():
            feats = tr.predict(imt / 255.0, torch.from_numpy(bm))
            if tag == "a":
                px = torch.stack([tr.clip_preprocess(imt / 255.0), tr.clip_preprocess((imt / 255.0)[:, 0:480, 0:640])])
                tok = model.visual.forward_features(px, norm=True)
                out["px_a_sub"] = px[:, :, ::7, ::7].numpy()
                out["tok_a_sub"] = tok[:, ::16].numpy()
                fm = tr.get_features_mask(torch.from_numpy(bm)) > 0
                out["fmask_a"] = np.packbits(fm.numpy(), axis=1)
        out[f"regions_{tag}"] = feats.numpy()
    with torch.no_grad():
        out["text"] = model.encode_text(torch.from_numpy(TEXT_TOKENS)).numpy()
    tok = rh.tokenizer(32)
    out["tokenizer_a_chair"] = tok(["a chair"]).numpy()
    np.savez_compressed(os.path.join(OUT, "encoder_tiny.npz"), **out)
    print("encoder_tiny.npz", {k: v.shape for k, v in out.items()})


ASSOC_CASES = [(0, 20000, 0.25), (3, 60000, 0.5), (7, 150000, 0.25)]


def gen_assoc():
    from ovo_b200 import synth
    rh.setup_paths()
    import ovo.utils.geometry_utils as gu
    K = synth.intrinsics()
    out = {}
    for fid, N, fv in ASSOC_CASES:
        c2w = synth.pose(fid); d = synth.depth_map(frame_id=fid)
        seg, _ = synth.grid_masks()
        xyz, ids, ins = synth.point_map(N, d, K, c2w, seed=fid, frac_visible=fv)
        tK, tc2w, td, txyz = map(torch.from_numpy, (K, c2w, d, xyz))
        corners = gu.compute_camera_frustum_corners(td, tc2w, tK)
        fmask = gu.compute_frustum_point_ids(txyz, corners, device="cpu")
        df = gu.depth_filter(td)
        midx, matches = gu.match_3d_points_to_2d_pixels(df, torch.linalg.inv(tc2w), txyz[fmask], tK, 0.05)
        seg_of_pt = np.full(N, -2, np.int16)
        seg_of_pt[fmask[midx].numpy()] = seg[matches[:, 1].numpy(), matches[:, 0].numpy()]
        out[f"seg_of_pt_{fid}"] = seg_of_pt
        out[f"depth_rejected_{fid}"] = np.packbits(df.numpy() == -1)
        out[f"corners_{fid}"] = corners.numpy()
    np.savez_compressed(os.path.join(OUT, "assoc.npz"), **out)
    print("assoc.npz", {k: v.shape for k, v in out.items()})


OVO_RUN = dict(n_points=40000, frac_visible=0.6, n_kf=4, pose_step=6, track_th=100)


def ovo_config(masks_dir):
    return {"segment_every": 1, "match_distance_th": 0.05, "track_th": OVO_RUN["track_th"], "depth_filter": True,
            "log": False, "kf_queue_delay": 1, "debug_info": False, "verbose": False,
            "sam": {"precomputed": True, "masks_base_path": masks_dir},
            "clip": {"embed_type": "TextRegion", "model_card": "PE-Core-L14-336", "k_top_views": 10000,
                     "fusion": "avg_pooling"}}


def ovo_inputs():
    """The synthetic replay every OVO end-to-end fixture / test uses."""
    from ovo_b200 import synth
    K = synth.intrinsics()
    d0 = synth.depth_map(frame_id=0)
    xyz, ids, ins = synth.point_map(OVO_RUN["n_points"], d0, K, synth.pose(0), seed=11, frac_visible=OVO_RUN["frac_visible"])
    frames = []
    for i in range(OVO_RUN["n_kf"]):
        fid = i * OVO_RUN["pose_step"]
        rows, cols = (6, 8) if i % 2 == 0 else (4, 5)       # masks change between keyframes -> merges and new ids
        seg, bm = synth.grid_masks(rows=rows, cols=cols)
        frames.append(dict(frame_id=fid, image=synth.rgb(seed=100 + i), depth=synth.depth_map(frame_id=fid),
                           c2w=synth.pose(fid), seg=seg, bm=bm))
    return K, xyz, ids, ins, frames


def gen_ovo():
    """Runs the reference OVO class itself (ovo/entities/ovo.py) over 4 keyframes."""
    rh.setup_paths()
    model, sd, cfg = build_reference_model()
    import ovo.utils.clip_utils as cu
    from torchvision.transforms import Resize, Normalize, CenterCrop, Compose
    import core.vision_encoder.transforms as transforms

    class TokTokenizer:          # queries are given as token-id rows in this fixture (tiny vocab)
        def __call__(self, phrase):
            return torch.from_numpy(QUERIES_TOK[int(phrase)][None])

    def fake_loader(model_card, ckpt_path=None):
        pre = transforms.get_image_transform(model.image_size)
        keep = [tf for tf in pre.transforms if isinstance(tf, (Resize, CenterCrop, Normalize))]
        return model, TokTokenizer(), Compose(keep)

    cu.load_perception_encoder = fake_loader
    from ovo.entities.ovo import OVO
    from ovo.entities.logger import Logger
    K, xyz, ids, ins, frames = ovo_inputs()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        mdir = os.path.join(tmp, "masks", "scene")
        os.makedirs(mdir)
        for f in frames:
            np.save(os.path.join(mdir, f"{f['frame_id']:04d}_seg_map_default.npy"), f["seg"])
            np.save(os.path.join(mdir, f"{f['frame_id']:04d}_bmap_default.npy"), f["bm"])
        logger = Logger(os.path.join(tmp, "log"), os.getpid(), False)
        ovo = OVO(ovo_config(os.path.join(tmp, "masks")), logger, scene_name="scene", cam_intrinsics=torch.from_numpy(K),
                  device="cpu")
        ovo.clip_generator.clip_dim = cfg.output_dim   # the reference hard-codes 1024 (clip_generator.py:41)
        pts, pids, pins = torch.from_numpy(xyz), torch.from_numpy(ids), torch.from_numpy(ins)
        for i, f in enumerate(frames):
            upd = ovo.detect_and_track_objects((f["frame_id"], f["image"], f["depth"], ()), (pts, pids, pins),
                                               torch.from_numpy(f["c2w"]))
            pins = upd
            out[f"ins_ids_{i}"] = upd.numpy().astype(np.int16)
            out[f"matched_ins_{i}"] = np.array(ovo.keyframes_queue[-1][0], np.int32)
            out[f"maps_area_{i}"] = ovo.keyframes_queue[-1][1].sum((1, 2)).numpy().astype(np.int32)
            ovo.compute_semantic_info()
        ovo.complete_semantic_info()
        keys = list(ovo.objects.keys())
        out["object_ids"] = np.array(keys, np.int32)
        out["object_clips"] = ovo.get_objs_clips().numpy()
        out["object_n_kfs"] = np.array([len(ovo.objects[k].kfs_ids) for k in keys], np.int32)
        out["query"] = ovo.query(["0", "1", "2"]).numpy()
        cls = ovo.classify_instances(["0", "1", "2"], template="{}", th=0.0)
        out["classes"], out["conf"] = cls["classes"], cls["conf"]
        cd = ovo.capture_dict(False)
        out["capture_keys"] = np.array(sorted(cd.keys()))
    np.savez_compressed(os.path.join(OUT, "ovo_run.npz"), **out)
    print("ovo_run.npz", {k: v.shape for k, v in out.items()})
    print("objects", keys, "n_kfs", out["object_n_kfs"])


MAPPER_FRAMES = [(fid, 3 * fid, 0.0123 * fid) for fid in range(5)]       # (depth frame, pose frame, yaw)


def gen_mapper():
    """Reference VanillaMapper.map (ovo/slam/vanilla_mapper.py:46-85) over 5 frames."""
    from ovo_b200 import synth
    rh.setup_paths()
    from ovo.slam.vanilla_mapper import VanillaMapper
    K = synth.intrinsics()
    vm = VanillaMapper({"device": "cpu", "mapping": {"k_pooling": 3}}, torch.from_numpy(K))
    out = {}
    for i, (fid, pf, yaw) in enumerate(MAPPER_FRAMES):
        d, c2w, img = synth.depth_map(frame_id=fid), synth.pose(pf, yaw=yaw), synth.rgb(seed=fid)
        n0 = vm.pcd.shape[0]
        vm.map([fid, img, d, c2w], torch.from_numpy(c2w))
        out[f"n_{i}"] = np.array(vm.pcd.shape[0] - n0)
        out[f"xyz_{i}"] = vm.pcd[n0:].numpy()[::37]
        out[f"col_{i}"] = vm.pcd_colors[n0:].numpy()[::37]
    np.savez_compressed(os.path.join(OUT, "mapper.npz"), **out)
    print("mapper.npz", {k: v.shape for k, v in out.items()})


def gen_masks():
    """Reference masks_update / mask2segmap (ovo/utils/segment_utils.py) on synthetic proposals."""
    rh.setup_paths()
    import ovo.utils.segment_utils as su
    from . import masks as OM
    out = {}
    for seed in (0, 1, 2):
        masks, iou, stab = OM.synth_masks(seed=seed)
        lst = [dict(segmentation=masks[i], predicted_iou=iou[i], stability_score=stab[i]) for i in range(len(masks))]
        kept, = su.masks_update(lst, iou_thr=0.8, score_thr=0.7, inner_thr=0.5)
        out[f"kept_{seed}"] = np.array([next(i for i in range(len(lst)) if lst[i] is k) for k in kept], np.int64)
        seg, bm = su.mask2segmap(kept, np.zeros(masks.shape[1:] + (3,)))
        out[f"seg_{seed}"] = seg.astype(np.int16)
        out[f"bm_{seed}"] = np.packbits(bm)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)
    print("masks.npz", {k: v.shape for k, v in out.items()})


HD80 = dict(width=160, layers=2, heads=2, mlp_width=640, output_dim=64, text_layers=0)   # head_dim 80, as the ViT-H/14-shaped encoder


def gen_encoder_hd80():
    """Reference PE VisionTransformer with head_dim 80 (BASELINE config 4's ViT-H/14 shape: 1280 / 16 heads): tokens of one image."""
    from ovo_b200.encoder import EncoderConfig, random_state_dict
    rh.setup_paths()
    from core.vision_encoder.config import PEConfig, PETextConfig
    import core.vision_encoder.pe as pe
    cfg = EncoderConfig(**HD80)
    vc = PEConfig(image_size=cfg.image_size, patch_size=cfg.patch_size, width=cfg.width, layers=cfg.layers, heads=cfg.heads,
                  mlp_ratio=cfg.mlp_width / cfg.width, pool_type="attn", output_dim=cfg.output_dim, use_cls_token=True,
                  attn_pooler_heads=cfg.heads)
    torch.manual_seed(123)
    model = pe.CLIP(vc, PETextConfig(context_length=32, width=128, heads=2, layers=1, output_dim=64, vocab_size=1024)).eval()
    missing, unexpected = model.load_state_dict(random_state_dict(cfg, seed=0, text=False), strict=False)
    assert not unexpected, unexpected
    torch.manual_seed(4)
    px = torch.randn(2, 3, 336, 336) * 0.5
    with torch.no_grad():
        tok = model.visual.forward_features(px, norm=True)
    out = {"tokens_sub": tok[:, ::9, ::4].numpy()}
    np.savez_compressed(os.path.join(OUT, "encoder_hd80.npz"), **out)
    print("encoder_hd80.npz", {k: v.shape for k, v in out.items()})


SAM_THR = dict(pred_iou_thresh=0.5, stability_score_thresh=0.5, box_nms_thresh=1.0)
SAM_AMG_HW = (240, 320)   # the AMG fixture uses a small frame so that the stored masks stay small   # random weights: the stock 0.8 / 0.95 would reject everything


SAM_OVO_SCORE_THR = 0.22


def sam_image(h=480, w=640, seed=5):
    """A blocky colour image with noise (something for the AA resize and the masks to bite on)."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (h // 40, w // 40, 3))
    img = np.kron(coarse, np.ones((40, 40, 1))) + rng.normal(0, 12, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def gen_sam():
    """Reference SAM-2 image path (SAM2Base + SAM2ImagePredictor + SAM2AutomaticMaskGenerator, tiny Hiera geometry,
    our seeded weights) and OVO's MaskGenerator.segment post-processing on one 640x480 image."""
    from ovo_b200.sam_config import tiny_sam_config, random_state_dict
    rh.setup_paths()
    import ovo.utils.segment_utils as su
    cfg = tiny_sam_config()
    sd = random_state_dict(cfg, seed=0)
    model = rh.build_sam2_reference(cfg, sd)
    amg = rh.build_sam2_amg(model, **SAM_THR)
    img = sam_image()
    out = {}
    with torch.no_grad():
        pred = amg.predictor
        px = pred._transforms(img)[None]
        out["px_sub"] = px[0, :, ::16, ::16].numpy()
        pred.set_image(img)
        emb = pred._features["image_embed"]
        s0, s1 = pred._features["high_res_feats"]
        out["embed_sub"] = emb[0, :, ::4, ::4].numpy()
        out["s0_sub"] = s0[0, :, ::16, ::16].numpy()
        out["s1_sub"] = s1[0, :, ::8, ::8].numpy()
        from sam2.utils.amg import build_point_grid
        pts = torch.as_tensor(build_point_grid(16) * np.array([[640, 480]]), dtype=torch.float32)
        in_pts = pred._transforms.transform_coords(pts, normalize=True, orig_hw=(480, 640))
        masks, iou, low = pred._predict(in_pts[:64, None, :], torch.ones(64, 1, dtype=torch.int), multimask_output=True,
                                        return_logits=True)
        out["iou64"] = iou.numpy()
        out["low_sub"] = low[:, :, ::16, ::16].numpy()
        out["low_full1"] = low[0, 0].numpy().astype(np.float16)
        out["mask_logit_sub"] = masks[:16, :, ::16, ::16].numpy()
        img = sam_image(*SAM_AMG_HW, seed=6)
        anns = amg.generate(img)
    out["n"] = np.array(len(anns))
    segs = np.stack([a["segmentation"] for a in anns])
    out["seg_bits_every8"] = np.packbits(segs[::8])
    out["seg_checksum"] = np.array([np.flatnonzero(m).sum() for m in segs], np.int64)   # sum of the set pixels' flat indices
    out["pred_iou"] = np.array([a["predicted_iou"] for a in anns], np.float32)
    out["stability"] = np.array([a["stability_score"] for a in anns], np.float32)
    out["bbox"] = np.array([a["bbox"] for a in anns], np.float32)
    out["area"] = np.array([a["area"] for a in anns], np.int64)
    out["points"] = np.array([a["point_coords"][0] for a in anns], np.float32)
    # OVO's second stage exactly as MaskGenerator.segment does it (mask_generator.py:113-119)
    kept, = su.masks_update(anns, iou_thr=0.8, score_thr=SAM_OVO_SCORE_THR, inner_thr=0.5)
    seg_map, bmaps = su.mask2segmap(kept, img)
    out["ovo_seg_map"] = seg_map.astype(np.int16)
    out["ovo_n"] = np.array(len(kept))
    out["ovo_bits"] = np.packbits(bmaps)
    np.savez_compressed(os.path.join(OUT, "sam_tiny.npz"), **out)
    print("sam_tiny.npz", {k: v.shape for k, v in out.items()}, "n", len(anns), "kept", len(kept))

CROP_POOL_HEADS = 2          # attn_pooler_heads of the tiny reference model (build_reference_model)
CROP_CASES = [("vanilla", 336), ("vanilla", 224), ("fixed_weights", 336), ("fixed_weights", 384), ("hovsg", 336),
              ("adaptive_weights", 336), ("concept_fusion", 336)]


def crop_masks(h=480, w=640):
    """Grid cells (touch every border: the margin box clamps), a disc (masked fill inside its box), a thin
    horizontal bar (padded square of the `vanilla` type), an L-shape."""
    from ovo_b200 import synth
    _, bm = synth.grid_masks(h, w, rows=2, cols=3)
    yy, xx = np.mgrid[0:h, 0:w]
    extra = np.zeros((3, h, w), bool)
    extra[0] = (yy - 200) ** 2 + (xx - 300) ** 2 < 90 ** 2
    extra[1, 100:112, 40:600] = True
    extra[2, 300:470, 500:530] = True
    extra[2, 440:470, 380:530] = True
    return np.concatenate([bm, extra])


def gen_crops():
    """The reference's crop-based descriptors: UNMODIFIED CLIPGenerator.extract_clip / segmap2segimg / fuse_clips
    (ovo/entities/clip_generator.py:125-158, ovo/utils/segment_utils.py:29-182, ovo/utils/clip_utils.py:21-48) with the
    un-vendored open_clip loader (clip_utils.py:51-88) returning the vendored pe.CLIP and its own transform."""
    from ovo_b200 import synth
    rh.setup_paths()
    model, sd, cfg = build_reference_model()
    import ovo.utils.clip_utils as cu
    import ovo.utils.segment_utils as su
    from torchvision.transforms import Resize, Normalize, CenterCrop, Compose
    import core.vision_encoder.transforms as transforms

    def fake_loader(model_card, use_half):
        pre = transforms.get_image_transform(model.image_size)
        keep = [tf for tf in pre.transforms if isinstance(tf, (Resize, CenterCrop, Normalize))]
        return model, None, Compose(keep), cfg.output_dim

    cu.load_clip_model = fake_loader
    from ovo.entities.clip_generator import CLIPGenerator
    img = synth.rgb(480, 640, seed=21)
    bm = crop_masks()
    imt = torch.from_numpy(img.transpose(2, 0, 1).copy())           # uint8 [3,H,W], what OVO._extract_clip passes (ovo.py:436)
    out = {}
    for et, res in CROP_CASES:
        gen = CLIPGenerator({"embed_type": et, "model_card": "PE-Core-L-14-336", "mask_res": res}, device="cpu")
        with torch.no_grad():
            out[f"{et}_{res}"] = gen.extract_clip(imt, torch.from_numpy(bm)).numpy()
            if et == "fixed_weights" and res == 336:
                out["return_all_336"] = gen.extract_clip(imt, torch.from_numpy(bm), return_all=True).numpy()
                out["encode_image_global"] = gen.encode_image(imt[None] / 255.).numpy()
    seg = su.segmap2segimg(torch.from_numpy(bm), imt, True, out_l=336)
    out["segimg_sub"] = seg[:, :, ::6, ::6].numpy()
    out["segimg_sum"] = seg.long().sum((2, 3)).numpy()
    segv = su.segmap2segimg(torch.from_numpy(bm), imt, False, out_l=224)
    out["segimg_vanilla_sub"] = segv[:, :, ::4, ::4].numpy()
    out["boxes_xywh"] = su.batched_box_xyxy_to_xywh(su.batched_mask_to_box(torch.from_numpy(bm))).numpy()
    np.savez_compressed(os.path.join(OUT, "crops.npz"), **out)
    print("crops.npz", {k: v.shape for k, v in out.items()})

def gen_labels():
    """The reference's own match_labels_to_vtx (ovo/utils/eval_utils.py:13-44; SciPy KD-tree + torch.mode)."""
    rh.setup_paths()
    import ovo.utils.eval_utils as eu
    from . import labels as OL
    out = {}
    for seed in (0, 1):
        pts, lab, vtx = OL.synth_scene(seed=seed)
        ml, masks, ids = eu.match_labels_to_vtx(torch.from_numpy(lab), torch.from_numpy(pts), torch.from_numpy(vtx))
        out[f"mesh_labels_{seed}"] = ml.numpy().astype(np.int16)
        out[f"ids_{seed}"] = ids.numpy()
        out[f"mask_sums_{seed}"] = masks.sum(1).numpy()
    pts, lab, vtx = OL.synth_scene(seed=0)
    ml, _, ids = eu.match_labels_to_vtx(torch.from_numpy(lab), torch.from_numpy(pts), torch.from_numpy(vtx), filter_unasigned=False)
    out["mesh_labels_0_unfiltered"] = ml.numpy().astype(np.int16)
    out["ids_0_unfiltered"] = ids.numpy()
    np.savez_compressed(os.path.join(OUT, "labels.npz"), **out)
    print("labels.npz", {k: v.shape for k, v in out.items()})

UPDATE_MAP = dict(th_centroid=1.5, th_cossim=0.5, th_points=0.5, kfs=[6, 18], drop_instance_rank=2)


def update_map_scenario(points_ins_ids: np.ndarray, object_ids):
    """Loop-closure input shared by the golden generator and the GPU test: the points of one instance lose their id (the SLAM
    back-end pruned them), keyframes 0 and 12 were culled."""
    ins = points_ins_ids.copy()
    drop = list(object_ids)[UPDATE_MAP["drop_instance_rank"]]
    ins[ins == drop] = -1
    return ins, UPDATE_MAP["kfs"], drop


def gen_update_map():
    """The reference's OVO.update_map (ovo/entities/ovo.py:366-424) + instance_utils.same_instance / fuse_instances after the
    4-keyframe replay of gen_ovo; Open3D's nearest-neighbour distance comes from oracle/shims/open3d."""
    rh.setup_paths()
    model, sd, cfg = build_reference_model()
    import ovo.utils.clip_utils as cu
    import ovo.utils.instance_utils as iu
    from torchvision.transforms import Resize, Normalize, CenterCrop, Compose
    import core.vision_encoder.transforms as transforms

    def fake_loader(model_card, ckpt_path=None):
        pre = transforms.get_image_transform(model.image_size)
        keep = [tf for tf in pre.transforms if isinstance(tf, (Resize, CenterCrop, Normalize))]
        return model, None, Compose(keep)

    cu.load_perception_encoder = fake_loader
    from ovo.entities.ovo import OVO
    from ovo.entities.logger import Logger
    K, xyz, ids, ins, frames = ovo_inputs()
    out, pairs = {}, []
    orig = iu.same_instance

    def logged(i1, i2, pc1, pc2, thc, ths, thp):
        r = orig(i1, i2, pc1, pc2, thc, ths, thp)
        cen = float(((pc1[1] - pc2[1]) ** 2).sum().sqrt())
        cos = float(torch.nn.functional.cosine_similarity(i1.clip_feature[0], i2.clip_feature[0], dim=0))
        pd = -1.0
        if cen <= thc and cos >= ths:
            import open3d as o3d
            a, b = o3d.geometry.PointCloud(), o3d.geometry.PointCloud()
            a.points, b.points = pc1[0].numpy(), pc2[0].numpy()
            pd = float((a.compute_point_cloud_distance(b) < thp).astype(float).mean())
        pairs.append((i1.id, i2.id, cen, cos, 1.0 if r else 0.0, pd))
        return r

    iu.same_instance = logged
    with tempfile.TemporaryDirectory() as tmp:
        mdir = os.path.join(tmp, "masks", "scene")
        os.makedirs(mdir)
        for f in frames:
            np.save(os.path.join(mdir, f"{f['frame_id']:04d}_seg_map_default.npy"), f["seg"])
            np.save(os.path.join(mdir, f"{f['frame_id']:04d}_bmap_default.npy"), f["bm"])
        logger = Logger(os.path.join(tmp, "log"), os.getpid(), False)
        config = ovo_config(os.path.join(tmp, "masks"))
        config.update({k: UPDATE_MAP[k] for k in ("th_centroid", "th_cossim", "th_points")})
        config["log"] = True         # keyframes["frame_id"] is only filled when logging (ovo.py:152-153): the culled-keyframe branch needs it
        torch.cuda.synchronize = lambda *a, **k: None     # the reference's profiler (ovo.py:101-118) synchronises CUDA; no GPU here
        ovo = OVO(config, logger, scene_name="scene", cam_intrinsics=torch.from_numpy(K), device="cpu")
        ovo.clip_generator.clip_dim = cfg.output_dim
        pts, pids, pins = torch.from_numpy(xyz), torch.from_numpy(ids), torch.from_numpy(ins)
        for f in frames:
            pins = ovo.detect_and_track_objects((f["frame_id"], f["image"], f["depth"], ()), (pts, pids, pins), torch.from_numpy(f["c2w"]))
            ovo.compute_semantic_info()
        before = list(ovo.objects.keys())
        ins_in, kfs, drop = update_map_scenario(pins.numpy(), before)
        upd = ovo.update_map((pts, pids, torch.from_numpy(ins_in)), kfs)
        out["ins_ids"] = upd.numpy().astype(np.int16)
        out["objects_before"] = np.array(before, np.int32)
        out["object_ids"] = np.array(list(ovo.objects.keys()), np.int32)
        out["object_clips"] = ovo.get_objs_clips().numpy()
        out["object_n_kfs"] = np.array([len(o.kfs_ids) for o in ovo.objects.values()], np.int32)
        out["object_n_points"] = np.array([len(o.points_ids) for o in ovo.objects.values()], np.int32)
        out["frame_id"] = np.array([str(x) for x in ovo.keyframes["frame_id"]])
        out["desc_keys"] = np.array(sorted(ovo.keyframes["ins_descriptors"].keys()), np.int32)
        out["pairs"] = np.array(pairs, np.float64)
    iu.same_instance = orig
    p = out["pairs"]
    # every decision must survive the GPU path's descriptor noise (cos +- 3e-3) and summation order (centroid +- 1e-3, p_dist +- 1e-3)
    def decide(cen, cos, pd):
        if cen > UPDATE_MAP["th_centroid"] or cos < UPDATE_MAP["th_cossim"]:
            return False
        return pd > 0.5 or (cos > 0.9 and pd > 0.2)
    fragile = [tuple(r) for r in p if r[5] >= 0 and len({decide(r[2] + a, r[3] + b, r[5] + c) for a in (-1e-3, 1e-3) for b in (-3e-3, 3e-3)
                                                         for c in (-1e-3, 1e-3)}) > 1]
    fragile += [tuple(r) for r in p if r[5] < 0 and (abs(r[2] - UPDATE_MAP["th_centroid"]) < 1e-3 and r[3] >= UPDATE_MAP["th_cossim"] - 3e-3
                                                    or abs(r[3] - UPDATE_MAP["th_cossim"]) < 3e-3 and r[2] <= UPDATE_MAP["th_centroid"] + 1e-3)]
    print("pairs", len(p), "fused", int(p[:, 4].sum()), "fragile decisions:", fragile)
    assert not fragile, "pick other thresholds: a decision sits on a threshold"
    np.savez_compressed(os.path.join(OUT, "update_map.npz"), **out)
    print("update_map.npz", {k: v.shape for k, v in out.items()})
    print("before", before, "after", out["object_ids"].tolist(), "dropped", drop)

MERGER_CFGS = {"per_channel": {"transformer": {"d_model": 64, "nhead": 8, "dim_feedforward": 128, "n_layers": 2},
                               "mlp": {"i_dim": 192, "h_dim": 256, "o_dim": 192, "n_layers": 2, "act_key": "leaky_relu"}},
               "per_clip": {"transformer": {"d_model": 64, "nhead": 4, "dim_feedforward": 64, "n_layers": 1},
                            "mlp": {"i_dim": 192, "h_dim": 128, "o_dim": 3, "n_layers": 1, "act_key": "leaky_relu"}}}


def merger_state_dict(cfg, seed=0):
    """Seeded weights of WeightsPredictorMerger (clips_merging.py:26-56) under its own state_dict keys."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    d, ff, sd = cfg["transformer"]["d_model"], cfg["transformer"]["dim_feedforward"], {}
    for l in range(cfg["transformer"]["n_layers"]):
        p = f"att_encoder.layers.{l}."
        sd[p + "self_attn.in_proj_weight"] = rn(3 * d, d, std=d ** -0.5); sd[p + "self_attn.in_proj_bias"] = rn(3 * d, std=0.02)
        sd[p + "self_attn.out_proj.weight"] = rn(d, d, std=d ** -0.5); sd[p + "self_attn.out_proj.bias"] = rn(d, std=0.02)
        sd[p + "linear1.weight"] = rn(ff, d, std=d ** -0.5); sd[p + "linear1.bias"] = rn(ff, std=0.02)
        sd[p + "linear2.weight"] = rn(d, ff, std=ff ** -0.5); sd[p + "linear2.bias"] = rn(d, std=0.02)
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = 1 + rn(d, std=0.05); sd[p + n + ".bias"] = rn(d, std=0.02)
    m = cfg["mlp"]
    dims = [m["i_dim"]] + [m["h_dim"]] * (m["n_layers"] + 1) + [m["o_dim"]]
    for j in range(len(dims) - 1):
        sd[f"mlp.{2 * j}.weight"] = rn(dims[j + 1], dims[j], std=2.0 * dims[j] ** -0.5); sd[f"mlp.{2 * j}.bias"] = rn(dims[j + 1], std=0.1)
    return sd


def gen_merger():
    """embed_type `learned`: the UNMODIFIED CLIPGenerator (learned branch, clip_generator.py:18-29,150-154) + WeightsPredictorMerger
    (clips_merging.py) loading a seeded checkpoint the way the reference loads its own (hparams.yaml + model.pt)."""
    import yaml
    from ovo_b200 import synth
    rh.setup_paths()
    model, sd, cfg = build_reference_model()
    import ovo.utils.clip_utils as cu
    from torchvision.transforms import Resize, Normalize, CenterCrop, Compose
    import core.vision_encoder.transforms as transforms

    def fake_loader(model_card, use_half):
        pre = transforms.get_image_transform(model.image_size)
        keep = [tf for tf in pre.transforms if isinstance(tf, (Resize, CenterCrop, Normalize))]
        return model, None, Compose(keep), cfg.output_dim

    cu.load_clip_model = fake_loader
    from ovo.entities.clip_generator import CLIPGenerator
    from ovo.entities.clips_merging import WeightsPredictorMerger
    img = synth.rgb(480, 640, seed=21)
    bm = crop_masks()
    imt = torch.from_numpy(img.transpose(2, 0, 1).copy())
    out = {}
    g = torch.Generator().manual_seed(9)
    clips = torch.nn.functional.normalize(torch.randn(11, 3, 64, generator=g), dim=-1)
    for name, mc in MERGER_CFGS.items():
        msd = merger_state_dict(mc, seed=5)
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "hparams.yaml"), "w") as f:
                yaml.safe_dump({"model": mc}, f)
            ref = WeightsPredictorMerger(mc).eval()
            ref.load_state_dict(msd)
            torch.save(ref.state_dict(), os.path.join(tmp, "model.pt"))
            gen = CLIPGenerator({"embed_type": "learned", "model_card": "PE-Core-L-14-336", "mask_res": 336,
                                 "weights_predictor_path": tmp}, device="cpu")
            with torch.no_grad():
                out[f"learned_{name}"] = gen.extract_clip(imt, torch.from_numpy(bm)).numpy()
                out[f"merge_{name}"] = ref(clips).numpy()
    np.savez_compressed(os.path.join(OUT, "merger.npz"), **out)
    print("merger.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not rh.available():
        sys.exit("reference not available: fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["encoder", "assoc", "ovo", "masks", "mapper", "sam", "encoder_hd80", "crops", "labels", "update_map", "merger"]
    for w in which:
        globals()["gen_" + w]()
