"""`MaskGenerator` with the reference's surface (ovo/entities/mask_generator.py:9-195).

Round-1 scope: the reference's own `sam.precomputed: True` seam (mask_generator.py:94-95,170-195) — masks are
read from `{frame_id:04d}_seg_map_default.npy` / `_bmap_default.npy` and uploaded.  The SAM-2 Hiera-L mask
proposal is not built yet; asking for it fails loudly instead of silently doing something
else."""
import os
from typing import Any, Dict, Tuple

import numpy as np
import torch


class MaskGenerator:
    def __init__(self, config: Dict[str, Any], scene_name: str | None = None, device="cuda") -> None:
        self.precomputed = config["precomputed"]
        self.config = config
        if scene_name:
            self.masks_path = os.path.join(config["masks_base_path"], scene_name)
        else:
            assert not config.get("precompute", False), "To precompute masks or use precomputed masks \"scene_name\" is required!"
            self.masks_path = ""
        self.nms_iou_th = config.get("nms_iou_th", 0.8)
        self.nms_score_th = config.get("nms_score_th", 0.7)
        self.nms_inner_th = config.get("nms_inner_th", 0.5)
        self.multi_crop = config.get("multi_crop", False)
        self.device = device
        self.mask_generator = None
        if not self.precomputed:
            raise NotImplementedError(
                "ovo_b200: on-line SAM-2 mask proposal is not built yet (SURVEY K22-K25); run with "
                "semantic.sam.precomputed: True and masks produced by the reference's MaskGenerator.precompute")

    def to(self, device: str) -> None:
        self.device = device

    def cpu(self) -> None:
        self.device = "cpu"

    def cuda(self) -> None:
        self.device = "cuda"

    def get_masks(self, image: np.ndarray, frame_id: int | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (seg_map [H,W] i32 with -1 = none, binary_maps [N,H,W] bool) on self.device (mask_generator.py:81-99)."""
        seg_map, binary_maps = self._load_masks(frame_id)
        return torch.from_numpy(seg_map).to(self.device), torch.from_numpy(binary_maps).to(self.device)

    def segment(self, image: np.ndarray):
        """mask_generator.py:101-120.  The proposal network (SAM-2 automatic mask generator) is not built here; when a
        `proposal_fn(image) -> list of {segmentation, predicted_iou, stability_score}` is attached (e.g. the
        reference's own SAM2AutomaticMaskGenerator.generate), its output goes through the native post-processing."""
        fn = getattr(self, "proposal_fn", None)
        if fn is None:
            raise NotImplementedError("ovo_b200: SAM-2 mask proposal is not built yet (attach MaskGenerator.proposal_fn)")
        masks = fn(image)
        if len(masks) == 0:
            return np.array([]), np.array([])
        seg, maps = self.postprocess(masks)
        return seg.cpu().numpy(), maps.cpu().numpy()

    def postprocess(self, masks):
        """masks_update + mask2segmap (segment_utils.py:173-259,12-27) on the device.
        masks: list of dicts with 'segmentation' [H,W] bool, 'predicted_iou', 'stability_score'.
        Returns (seg_map [H,W] i32, binary_maps [M,H,W] bool) as device tensors."""
        from .map import SemanticMap
        if getattr(self, "_sm", None) is None:
            self._sm = SemanticMap(self.device if "cuda" in str(self.device) else "cuda")
        seg = torch.from_numpy(np.stack([m["segmentation"] for m in masks]))
        iou = torch.tensor([float(m["predicted_iou"]) for m in masks], dtype=torch.float32)
        stab = torch.tensor([float(m["stability_score"]) for m in masks], dtype=torch.float32)
        seg_d = seg.to(self._sm.device)
        keep = self._sm.mask_nms(seg_d, stab * iou, self.nms_iou_th, self.nms_score_th, self.nms_inner_th)
        idx = torch.nonzero(keep).flatten()
        if idx.numel() == 0:
            return torch.full(seg.shape[1:], -1, dtype=torch.int32, device=self._sm.device), seg_d[:0]
        seg_map, maps, _ = self._sm.mask2segmap(seg_d[idx], stab.to(self._sm.device)[idx])
        return seg_map, maps

    def precompute(self, dataset, segment_every: int) -> None:
        """With every mask already on disk this is the reference's no-op path (mask_generator.py:141-152)."""
        for frame_id in range(len(dataset)):
            if frame_id % segment_every:
                continue
            a = os.path.join(self.masks_path, f"{frame_id:04d}_seg_map_default.npy")
            b = os.path.join(self.masks_path, f"{frame_id:04d}_bmap_default.npy")
            if not (os.path.exists(a) and os.path.exists(b)):
                self.segment(dataset[frame_id][1])
        self.precomputed = True

    def _load_masks(self, frame_id: int) -> Tuple[np.ndarray, np.ndarray]:
        map_path = os.path.join(self.masks_path, f"{frame_id:04d}_seg_map_default.npy")
        if not os.path.exists(map_path):
            print(f"No precomputed mask for frame {frame_id}")
            return np.array([]), np.array([])
        seg_map = np.load(map_path)
        bin_path = os.path.join(self.masks_path, f"{frame_id:04d}_bmap_default.npy")
        if os.path.exists(bin_path):
            binary_maps = np.load(bin_path)
        else:   # rebuild the binary maps from the seg map (mask_generator.py:186-188)
            idxs = np.arange(seg_map.max() + 1)
            binary_maps = seg_map[None] == idxs[:, None, None]
        return seg_map, binary_maps
