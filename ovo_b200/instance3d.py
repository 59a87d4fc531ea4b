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

    @property
    def clip_feature_kf(self):
        """Index of the view a medoid fusion picked (instance3d.py:185-187).  The fusion kernel leaves it on the device; it is
        read back on first use (export) instead of blocking the host behind the encoder at every keyframe."""
        v = self._clip_feature_kf
        if isinstance(v, tuple):           # (device tensor of one launch's picks, position)
            v = self._clip_feature_kf = int(v[0][v[1]].item())
        return v

    @clip_feature_kf.setter
    def clip_feature_kf(self, v) -> None:
        self._clip_feature_kf = v

    def __init__(self, id: int, kf_id: int | None = None, points_ids: List[int] | None = None, mask_area: int = 0):
        self.id = id
        self.clip_feature = None           # torch view of the bank row: [D] (one view) or [1,D] (fused), instance3d.py:184-187
        self.clip_feature_kf = None
        self.kfs_ids: List[int] = []
        self.points_ids: List[int] = []
        self.top_kf: List[tuple] = []      # min-heap of (area, kf_id)
        self.to_update = False
        self.bank_row = -1                 # row of the device instance bank (set by OVO)
        # host-side indices: membership tests in O(1) (the reference scans its lists, instance3d.py:105-155: per-keyframe cost
        # grows with the length of the stream) and the bookkeeping of the incremental avg_pooling fusion (ovo_b200/ovo.py)
        self._kfs_set = set()
        self._top_area: Dict[int, int] = {}
        self.pending_rows: List[int] = []  # descriptor-store rows computed for this instance but not yet in the fused feature
        self.n_fused = 0                   # views behind the current bank row
        self.evicted = False               # a view left the top-k heap or changed: the next fusion must re-read every view
        self._desc_epoch = 0               # OVO._desc_epoch at the last full fusion
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
        if len(self._kfs_set) != len(self.kfs_ids):  # the list was assigned from outside (restore, merges)
            self._kfs_set = set(self.kfs_ids)
        if kf_id not in self._kfs_set:
            self._kfs_set.add(kf_id)
            self.kfs_ids.append(kf_id)

    def _sync_top_index(self) -> None:
        if len(self._top_area) != len(self.top_kf):  # the heap was assigned from outside
            self._top_area = {k: a for a, k in self.top_kf}

    def idx_in_top_kf(self, kf_id: int) -> int:
        self._sync_top_index()
        if kf_id not in self._top_area:
            return -1
        for i, (_, k) in enumerate(self.top_kf):
            if k == kf_id:
                return i
        return -1

    def is_top_kf(self, kf_id: int) -> bool:
        self._sync_top_index()
        return kf_id in self._top_area

    def add_top_kf(self, kf_id: int, area: int) -> None:
        self._sync_top_index()
        if kf_id in self._top_area:                  # known keyframe: keep the larger area
            if area > self._top_area[kf_id]:
                i = self.idx_in_top_kf(kf_id)
                self.top_kf[i] = (area, kf_id)
                self._top_area[kf_id] = area
                heapq.heapify(self.top_kf)
                self.to_update = True
                self.evicted = True                  # the order of the views changed
            return
        if len(self.top_kf) < self.n_top_kf:
            heapq.heappush(self.top_kf, (area, kf_id))
            self._top_area[kf_id] = area
            self.to_update = True
        else:
            dropped = heapq.heappushpop(self.top_kf, (area, kf_id))
            if self.n_top_kf > 0:
                self._top_area[kf_id] = area
                self._top_area.pop(dropped[1], None)
                self.evicted = True
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
        self.n_fused, self.evicted, self.pending_rows = len(views), False, []
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
        self.evicted = True                          # the next fusion re-reads every view
        if debug_info:
            self.kfs_ids = obj_dict[f"ins3d_{self.id}_keyframes_ids"].tolist()
            self.points_ids = obj_dict[f"ins3d_{self.id}_points_ids"].tolist()
            if obj_dict.get(f"ins3d_{self.id}_top_kfs", None) is not None:
                self.top_kf = [(a, k) for a, k in obj_dict[f"ins3d_{self.id}_top_kfs"]]
        self._kfs_set = set(self.kfs_ids)
        self._top_area = {k: a for a, k in self.top_kf}

    def purge_points_ids(self, purge_ids: List[int]) -> None:
        drop = set(purge_ids)
        self.points_ids = [p for p in self.points_ids if p not in drop]
