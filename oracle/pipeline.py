"""TEST INFRASTRUCTURE — the OVO keyframe loop restated on top of oracle/encoder.py + oracle/fusion.py
(ovo/entities/ovo.py:121-166,326-364,440-527; ovo/entities/instance3d.py:157-189) with the default
configuration (fusion avg_pooling, k_top_views = all).  Pinned by tests/golden/ovo_run.npz, which was
produced by the reference's own OVO class."""
import numpy as np
import torch

from . import encoder as OE
from . import fusion as OF


class OracleOVO:
    def __init__(self, W: dict, cfg: OE.VitCfg, K: np.ndarray, match_th=0.05, track_th=100, depth_filter=True,
                 kf_queue_delay=1, encode_fn=None):
        self.W, self.cfg, self.K = W, cfg, K
        # descriptor rule: TextRegion by default; the crop-based types pass oracle.crops.extract_clip (clip_generator.py:125-158)
        self.encode_fn = encode_fn or (lambda image, masks: OE.encode_regions(image, masks, W, cfg))
        self.match_th, self.track_th, self.use_df, self.delay = match_th, track_th, depth_filter, kf_queue_delay
        self.next_ins_id, self.kf_id = 0, 0
        self.queue = []
        self.kf_desc = {}            # kf -> {ins: feat}
        self.objects = {}            # ins -> dict(kfs=[...], clip=None)   (insertion order = creation order)
        self.log = []                # per keyframe: (matched_ins_ids, areas)

    def detect_and_track(self, image, depth, c2w, seg, bmaps, xyz, ins_ids, rgb_depth_ratio=()):
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        seg_of_pt, _ = OF.associate(xyz, ins_ids, depth, seg, c2w, w2c, self.K, self.match_th, self.use_df, rgb_depth_ratio)
        ins_new, rows, self.next_ins_id = OF.track(ins_ids, seg_of_pt, seg, self.track_th, self.next_ins_id)
        for r in rows:
            if r["ins_id"] > -1:
                ob = self.objects.setdefault(r["ins_id"], dict(kfs=[], clip=None, to_update=False))
                if self.kf_id not in ob["kfs"]:
                    ob["kfs"].append(self.kf_id)
                # Instance3D.add_top_kf (instance3d.py:105-137): a new keyframe entering the top-k heap (k = 10000,
                # i.e. always) raises to_update; it is only lowered by update_clip.
                ob["to_update"] = True
        order, fused, mask_row = OF.fuse_masks(bmaps, rows)
        self.queue.append((order, fused, image, self.kf_id, seg_of_pt, mask_row))
        self.log.append((list(order), fused.sum((1, 2)).astype(np.int32), rows, seg_of_pt, mask_row))
        self.kf_id += 1
        return ins_new

    def compute_semantic_info(self, flush=False):
        while len(self.queue) > (0 if flush else self.delay):
            order, fused, image, kf, _, _ = self.queue.pop(0)
            if len(order) == 0:
                if not flush:
                    break
                continue
            with torch.no_grad():
                feats = self.encode_fn(image, fused)
            self.kf_desc[kf] = {ins: feats[i] for i, ins in enumerate(order)}
            for ins in order:                                   # Instance3D.update_clip, avg_pooling
                # instance3d.py:157-189: recomputed only while to_update is set.  Because CLIP runs kf_queue_delay
                # keyframes late, the flag is usually already lowered when the LAST keyframes are flushed by
                # complete_semantic_info, so their descriptors never enter the fused feature (reference quirk).
                if not self.objects[ins]["to_update"]:
                    continue
                clips = [self.kf_desc[k][ins] for k in self.objects[ins]["kfs"] if k in self.kf_desc and ins in self.kf_desc[k]]
                if len(clips) > 0:
                    self.objects[ins]["to_update"] = False
                if len(clips) == 1:
                    self.objects[ins]["clip"] = clips[0]
                elif len(clips) > 1:
                    self.objects[ins]["clip"] = torch.stack(clips).mean(0)
            if not flush:
                break

    def bank(self) -> torch.Tensor:
        return torch.stack([o["clip"] for o in self.objects.values()])

    def query(self, per_query_tokens) -> torch.Tensor:
        with torch.no_grad():
            txt = OE.text_bank(per_query_tokens, self.W, self.cfg)
        return OE.cosine_query(self.bank(), txt)
