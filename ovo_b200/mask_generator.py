"""`MaskGenerator` with the reference's surface (ovo/entities/mask_generator.py:9-195).

Two sources of masks, as in the reference: the `sam.precomputed: True` seam (mask_generator.py:94-95,170-195) — masks
read from `{frame_id:04d}_seg_map_default.npy` / `_bmap_default.npy` and uploaded — and the on-line SAM-2 automatic mask
generator (mask_generator.py:37-53,102-120) running on the device through `ovo_b200.sam.Sam2` (ovo_sam_generate).

Weights: the reference loads `sam{version}_{hiera_large}.pt` from `sam_ckpt_path` (segment_utils.py:269-295).  Here the
state_dict comes from `config["sam_state_dict"]` (a dict with the reference's key names), from that checkpoint file if it
exists, or — when `config["sam_random_init"]` is set (benchmarks, tests: no checkpoints are available offline) — from
seeded random weights of the configured geometry."""
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
        if (self.precomputed or config.get("precompute", False)) and os.path.isdir(self.masks_path):
            pass      # mask_generator.py:31-33: masks already on disk, no network needed
        elif not self.precomputed:
            self.load_mask_generator(config)

    def load_mask_generator(self, config: Dict[str, Any]) -> None:
        """mask_generator.py:37-53 + segment_utils.load_sam (:269-309): SAM-2.1 Hiera-L + the AMG thresholds OVO wires."""
        from .sam import Sam2
        from .sam_config import SamConfig, random_state_dict
        if config.get("sam_version", "2.1") == "":
            raise NotImplementedError("ovo_b200: SAM-1 (segment_anything, un-vendored in the reference) is not built; use sam_version 2.1")
        cfg = config.get("sam_config") or SamConfig()
        sd = config.get("sam_state_dict")
        if sd is None:
            ckpt = os.path.join(config.get("sam_ckpt_path", ""), f"sam{config.get('sam_version', '2.1')}_hiera_large.pt")
            if os.path.exists(ckpt):
                sd = torch.load(ckpt, map_location="cpu", weights_only=True)
                sd = sd.get("model", sd)
            elif config.get("sam_random_init", False):
                sd = random_state_dict(cfg, seed=int(config.get("sam_seed", 0)))
            else:
                raise FileNotFoundError(f"ovo_b200: SAM-2 checkpoint {ckpt} not found (pass sam_state_dict, or sam_random_init for benchmarks)")
        n = int(config.get("points_per_side", 32))
        self.mask_generator = Sam2(cfg, sd, max_h=int(config.get("max_h", 968)), max_w=int(config.get("max_w", 1296)),
                                   max_prompts=n * n, device=self.device, max_batch=max(1, int(config.get("batch_frames", 4))))
        self.amg_params = Sam2.amg_params(points_per_side=n, pred_iou_thresh=config.get("nms_iou_th", 0.8),
                                          stability_score_thresh=config.get("stability_score_th", 0.95),
                                          box_nms_thresh=config.get("box_nms_thresh", 0.7),
                                          nms_iou_th=self.nms_iou_th, nms_score_th=self.nms_score_th, nms_inner_th=self.nms_inner_th)

    def to(self, device: str) -> None:
        self.device = device

    def cpu(self) -> None:
        self.device = "cpu"

    def cuda(self) -> None:
        self.device = "cuda"

    def get_masks(self, image: np.ndarray, frame_id: int | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (seg_map [H,W] i32 with -1 = none, binary_maps [N,H,W] bool) on self.device (mask_generator.py:81-99)."""
        if self.precomputed:
            seg_map, binary_maps = self._load_masks(frame_id)
            return torch.from_numpy(seg_map).to(self.device), torch.from_numpy(binary_maps).to(self.device)
        return self.segment_device(image)

    def segment_device(self, image) -> Tuple[torch.Tensor, torch.Tensor]:
        """segment() without the round trip through host numpy: image HxWx3 uint8 (numpy or device tensor) ->
        (seg_map [H,W] i32, binary_maps [M,H,W] bool) on the device; empty tensors when no mask survives."""
        if self.mask_generator is None:
            raise RuntimeError("ovo_b200: MaskGenerator has no SAM-2 loaded (precomputed masks were configured)")
        img = image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))
        if img.dtype != torch.uint8:      # the reference's ToTensor only rescales uint8 input (transforms.py:37-39)
            raise TypeError("ovo_b200: SAM-2 expects a uint8 HxWx3 image")
        seg, maps = self.mask_generator.generate(img, self.amg_params, max_masks=int(self.config.get("max_masks", 256)))
        if maps.shape[0] == 0:
            return torch.zeros(0, device=maps.device), torch.zeros(0, device=maps.device)
        return seg, maps

    def segment(self, image: np.ndarray):
        """mask_generator.py:101-120 -> (seg_map, binary_maps) as numpy.  When a `proposal_fn(image) -> list of
        {segmentation, predicted_iou, stability_score}` is attached (e.g. masks from another proposal network), its output
        goes through the native post-processing instead of the built-in SAM-2."""
        fn = getattr(self, "proposal_fn", None)
        if fn is None:
            seg, maps = self.segment_device(image)
            return seg.cpu().numpy(), maps.cpu().numpy()
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
        """mask_generator.py:122-152: segment every `segment_every`-th frame that has no masks on disk yet and save them as
        `{frame_id:04d}_seg_map_default.npy` / `_bmap_default.npy`.  Frames are sent through SAM-2 `sam.batch_frames` at a
        time (default 4): one trunk pass per batch (the results equal frame-by-frame calls, tests/test_gpu_sam.py)."""
        print("Precomputing segmentation masks.")
        os.makedirs(self.masks_path, exist_ok=True)
        todo = []
        for frame_id in range(len(dataset)):
            if frame_id % segment_every:
                continue
            a = os.path.join(self.masks_path, f"{frame_id:04d}_seg_map_default.npy")
            b = os.path.join(self.masks_path, f"{frame_id:04d}_bmap_default.npy")
            if os.path.exists(a) and os.path.exists(b):
                print(f"Frame {frame_id} already compute. Skipping ...")
            else:
                todo.append(frame_id)
        if todo and self.mask_generator is None and getattr(self, "proposal_fn", None) is None:
            self.load_mask_generator(self.config)
        B = max(1, min(int(self.config.get("batch_frames", 4)), getattr(self.mask_generator, "max_batch", 1)))
        i = 0
        while i < len(todo):
            chunk = todo[i: i + B]
            images = [np.ascontiguousarray(dataset[f][1]) for f in chunk]
            same = all(im.shape == images[0].shape and im.dtype == np.uint8 for im in images)
            if len(chunk) > 1 and same and getattr(self, "proposal_fn", None) is None:
                outs = self.mask_generator.generate_batch(torch.from_numpy(np.stack(images)), self.amg_params,
                                                          max_masks=int(self.config.get("max_masks", 256)))
                for f, (seg, maps) in zip(chunk, outs):
                    if maps.shape[0] == 0:
                        self._save_masks(np.array([]), np.array([]), f)
                    else:
                        self._save_masks(seg.cpu().numpy(), maps.cpu().numpy(), f)
            else:
                for f, im in zip(chunk, images):
                    seg, maps = self.segment(im)
                    self._save_masks(seg, maps, f)
            i += len(chunk)
        self.precomputed = True

    def _save_masks(self, seg_map: np.ndarray, binary_maps: np.ndarray, frame_id: int) -> None:
        """mask_generator.py:154-168."""
        np.save(os.path.join(self.masks_path, f"{frame_id:04d}_seg_map_default"), seg_map)
        np.save(os.path.join(self.masks_path, f"{frame_id:04d}_bmap_default"), binary_maps)

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
