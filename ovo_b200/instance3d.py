"""Host bookkeeping of a 3D instance — same public surface as the reference's
`ovo/entities/instance3d.py:28-252` (attributes id, clip_feature, clip_feature_kf, kfs_ids, points_ids,
top_kf, to_update; methods update, add_points_ids, add_keyframes, add_top_kf, is_top_kf, idx_in_top_kf,
update_clip, export, restore, purge_points_ids; class attributes n_top_kf, mv_fusion).

The descriptor arithmetic does not live here: per-keyframe descriptors are rows of a device-resident store
and the fused descriptor is a row of the device-resident instance bank; `OVO` batches the fusion of all
instances touched by a keyframe into ONE `ovo_fuse_views` launch (see ovo_b200/ovo.py).  This class only
decides WHICH views are fused, exactly like the reference (top-k heap by mask area, `to_update` flag).
"""
import heapq
from typing import Any, Dict, List

import numpy as np

FUSION_MODES = {"avg_pooling": 0, "l1_medoid": 1, "cossim_medoid": 2}


class Instance3D:
    n_top_kf: int = 0
    mv_fusion: str = "l1_medoid"          # reference default (instance3d.py:51)

    def __init__(self, id: int, kf_id: int | None = None, points_ids: List[int] | None = None, mask_area: int = 0):
        self.id = id
        self.clip_feature = None           # torch view of the bank row: [D] (one view) or [1,D] (fused), instance3d.py:184-187
        self.clip_feature_kf = None
        self.kfs_ids: List[int] = []
        self.points_ids: List[int] = []
        self.top_kf: List[tuple] = []      # min-heap of (area, kf_id)
        self.to_update = False
        self.bank_row = -1                 # row of the device instance bank (set by OVO)
        if kf_id is not None:
            self.update(points_ids or [], kf_id, mask_area)

    @staticmethod
    def set_fusion(fusion: str, ckpt=None) -> None:
        if fusion == "camfusion":
            raise NotImplementedError("CAMFusion loading function not implemented yet.")   # clip_utils.py:114-115
        if fusion not in FUSION_MODES:
            raise NotImplementedError()
        Instance3D.mv_fusion = fusion

    # ---------------------------------------------------------------- bookkeeping (instance3d.py:77-155)
    def update(self, points_ids: List[int], kf_id: int, area: int) -> None:
        self.add_keyframes(kf_id)
        self.add_points_ids(points_ids)
        self.add_top_kf(kf_id, area)

    def add_points_ids(self, points_ids: List[int]) -> None:
        self.points_ids.extend(points_ids)

    def add_keyframes(self, kf_id: int) -> None:
        if kf_id not in self.kfs_ids:
            self.kfs_ids.append(kf_id)

    def idx_in_top_kf(self, kf_id: int) -> int:
        for i, (_, k) in enumerate(self.top_kf):
            if k == kf_id:
                return i
        return -1

    def is_top_kf(self, kf_id: int) -> bool:
        return self.idx_in_top_kf(kf_id) > -1

    def add_top_kf(self, kf_id: int, area: int) -> None:
        i = self.idx_in_top_kf(kf_id)
        if i > -1:                                   # known keyframe: keep the larger area
            if area > self.top_kf[i][0]:
                self.top_kf[i] = (area, kf_id)
                heapq.heapify(self.top_kf)
                self.to_update = True
            return
        if len(self.top_kf) < self.n_top_kf:
            heapq.heappush(self.top_kf, (area, kf_id))
            self.to_update = True
        else:
            dropped = heapq.heappushpop(self.top_kf, (area, kf_id))
            if self.n_top_kf <= 0 or dropped[1] != kf_id:
                self.to_update = True

    # ---------------------------------------------------------------- view selection (instance3d.py:157-189)
    def views_to_fuse(self, keyframes_clips: Dict[int, Dict[int, Any]], force_update: bool = False):
        """Returns the list of per-keyframe descriptor handles to fuse now, or None if nothing is to be done.
        Mirrors update_clip's selection: top-k keyframes by area (descending) when n_top_kf > 0, else every
        keyframe the instance was seen in; keyframes without descriptors yet are skipped."""
        if not (self.to_update or force_update):
            return None
        kfs = [kf for _, kf in heapq.nlargest(self.n_top_kf, self.top_kf)] if self.n_top_kf > 0 else self.kfs_ids
        views = []
        for kf in kfs:
            kf_clips = keyframes_clips.get(kf)
            if kf_clips is not None:
                views.append(kf_clips[self.id])
        if len(views) == 0:
            return None
        self.to_update = False
        return views

    def update_clip(self, keyframes_clips: Dict[int, Dict[int, Any]], force_update: bool = False, fuser=None) -> None:
        """Single-instance form kept for API compatibility; `fuser(instance, views)` performs the device fusion
        (OVO passes its batched fuser)."""
        views = self.views_to_fuse(keyframes_clips, force_update)
        if views is not None:
            if fuser is None:
                raise RuntimeError("Instance3D.update_clip needs the OVO fuser (descriptors live on the device)")
            fuser([(self, views)])

    # ---------------------------------------------------------------- (de)serialisation (instance3d.py:191-227)
    def export(self, debug_info: bool = False) -> Dict[str, Any]:
        d = {f"ins3d_{self.id}_clip_feature": self.clip_feature,
             f"ins3d_{self.id}_clip_feature_kf": self.clip_feature_kf}
        if debug_info:
            d.update({f"ins3d_{self.id}_keyframes_ids": np.array(self.kfs_ids),
                      f"ins3d_{self.id}_points_ids": np.array(self.points_ids),
                      f"ins3d_{self.id}_top_kfs": np.array(self.top_kf)})
        return d

    def restore(self, obj_dict: Dict[str, Any], debug_info: bool) -> None:
        self.clip_feature = obj_dict[f"ins3d_{self.id}_clip_feature"]
        self.clip_feature_kf = obj_dict.get(f"ins3d_{self.id}_clip_feature_kf", None)
        self.to_update = self.clip_feature is None
        if debug_info:
            self.kfs_ids = obj_dict[f"ins3d_{self.id}_keyframes_ids"].tolist()
            self.points_ids = obj_dict[f"ins3d_{self.id}_points_ids"].tolist()
            if obj_dict.get(f"ins3d_{self.id}_top_kfs", None) is not None:
                self.top_kf = [(a, k) for a, k in obj_dict[f"ins3d_{self.id}_top_kfs"]]

    def purge_points_ids(self, purge_ids: List[int]) -> None:
        drop = set(purge_ids)
        self.points_ids = [p for p in self.points_ids if p not in drop]
