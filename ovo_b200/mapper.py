"""`PointMapper` — the map producer with the surface of the reference's `VanillaMapper`
(ovo/slam/vanilla_mapper.py:7-136: track_camera, map, get_c2w, get_map, get_kfs, get_map_dict/set_map_dict,
get_cam_dict/set_cam_dict, update_pcd_obj_ids, get_pcd_colors), SURVEY §8f rank 3.

Differences underneath: one fused device pass decides which depth pixels are new (cull + project + depth test of
the existing map, 3x3 erosion, stride-2 sampling) and appends them into PRE-RESERVED buffers that grow
geometrically — the reference re-allocates the whole map with torch.vstack on every mapped frame
(vanilla_mapper.py:81-85), which is what breaks the 0 -> 8M point streaming configuration."""
import ctypes as C
from typing import Any, Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .map import SemanticMap


class PointMapper:
    def __init__(self, config: dict, cam_intrinsics: torch.Tensor, semmap: SemanticMap | None = None, capacity: int = 1 << 20) -> None:
        self.cam_intrinsics = cam_intrinsics
        self.config = config
        self.device = torch.device(config.get("device", "cuda"))
        if self.device.type != "cuda":
            raise RuntimeError("ovo_b200.PointMapper needs a CUDA device (no CPU fallback)")
        self.match_distance_th = 0.03
        self.max_id = 0
        self.estimated_c2ws: Dict[int, torch.Tensor] = {}
        self.kfs: Dict[int, Dict[str, Any]] = {}
        self.map_updated = False
        self.k_pool = int(config["mapping"].get("k_pooling", 3))
        self.downscale = int(config["mapping"].get("downscale_res", 2))
        self.semmap = semmap or SemanticMap(self.device)
        self.n = 0
        self._alloc(int(config["mapping"].get("reserve_points", capacity)))

    def _alloc(self, cap: int) -> None:
        old = getattr(self, "_xyz", None)
        xyz = torch.empty(cap, 3, device=self.device, dtype=torch.float32)
        ids = torch.empty(cap, device=self.device, dtype=torch.int32)
        obj = torch.full((cap,), -1, device=self.device, dtype=torch.int32)
        col = torch.empty(cap, 3, device=self.device, dtype=torch.uint8)
        if old is not None:
            xyz[: self.n], ids[: self.n], obj[: self.n], col[: self.n] = self._xyz[: self.n], self._ids[: self.n], self._obj[: self.n], self._col[: self.n]
        self._xyz, self._ids, self._obj, self._col, self.capacity = xyz, ids, obj, col, cap

    # the reference exposes these as attributes
    @property
    def pcd(self): return self._xyz[: self.n]
    @property
    def pcd_ids(self): return self._ids[: self.n, None]
    @property
    def pcd_obj_ids(self): return self._obj[: self.n, None]
    @property
    def pcd_colors(self): return self._col[: self.n]

    def track_camera(self, frame_data: List[Any]) -> None:
        frame_id, c2w = frame_data[0], frame_data[3]
        if np.isinf(c2w).sum() > 0 or np.isnan(c2w).sum() > 0:
            return
        self.estimated_c2ws[frame_id] = torch.from_numpy(np.asarray(c2w, np.float32))

    def map(self, frame_data: List[Any], c2w) -> int:
        """vanilla_mapper.py:46-85.  Returns the number of points added."""
        image, depth = frame_data[1], frame_data[2]
        h, w = depth.shape
        need = self.n + ((h + self.downscale - 1) // self.downscale) * ((w + self.downscale - 1) // self.downscale)
        if need > self.capacity:
            self._alloc(max(need, 2 * self.capacity))
        depth_d = torch.as_tensor(np.ascontiguousarray(depth, dtype=np.float32)).to(self.device, non_blocking=True)
        rgb_d = torch.as_tensor(np.ascontiguousarray(image, dtype=np.uint8)).to(self.device, non_blocking=True)
        c2w_np = (c2w.detach().float().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w, np.float32)).reshape(4, 4)
        w2c_np = torch.linalg.inv(torch.from_numpy(c2w_np)).numpy()
        K_np = (self.cam_intrinsics.detach().float().cpu().numpy() if torch.is_tensor(self.cam_intrinsics) else np.asarray(self.cam_intrinsics, np.float32)).reshape(3, 3)
        f16, f9 = C.c_float * 16, C.c_float * 9
        n_new = C.c_int(0)
        check(self.semmap.lib.ovo_map_integrate(self.semmap.handle, ptr(self._xyz), ptr(self._ids), ptr(self._obj), ptr(self._col),
                                                self.n, self.capacity, ptr(depth_d), ptr(rgb_d), h, w, f16(*c2w_np.reshape(-1).tolist()),
                                                f16(*w2c_np.reshape(-1).tolist()), f9(*K_np.reshape(-1).tolist()),
                                                float(self.match_distance_th), self.downscale, self.k_pool, self.max_id, C.byref(n_new),
                                                stream_ptr()), "ovo_map_integrate")
        self.n += n_new.value
        self.max_id += n_new.value
        return n_new.value

    def get_c2w(self, frame_id: int):
        return self.estimated_c2ws.get(frame_id, None)

    def cam_to_cpu(self, frame_id: int) -> None:
        pass

    def get_map(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return self.pcd, self.pcd_ids, self._obj[: self.n]

    def get_kfs(self) -> Dict[int, Dict[str, Any]]:
        return self.kfs

    def update_pcd_obj_ids(self, pcd_objs_ids: torch.Tensor) -> None:
        self._obj[: self.n] = pcd_objs_ids.reshape(-1).to(self.device, torch.int32)

    def get_pcd_colors(self) -> np.ndarray:
        return self.pcd_colors.cpu().numpy()

    def get_map_dict(self) -> Dict[str, Any]:
        return {"xyz": self.pcd.clone().cpu(), "obj_ids": self.pcd_obj_ids.clone().cpu(), "ids": self.pcd_ids.clone().cpu(),
                "max_id": self.max_id, "color": self.pcd_colors.clone().cpu()}

    def set_map_dict(self, map_dict: Dict[str, Any]) -> None:
        d = map_dict
        n = d["xyz"].shape[0]
        if n > self.capacity:
            self._alloc(2 * n)
        self.n = n
        self._xyz[:n] = d["xyz"].to(self.device)
        self._obj[:n] = d["obj_ids"].reshape(-1).to(self.device, torch.int32)
        self._ids[:n] = d["ids"].reshape(-1).to(self.device, torch.int32)
        self._col[:n] = d["color"].to(self.device)
        self.max_id = d["max_id"]

    def get_cam_dict(self) -> dict:
        return {k: v.cpu().numpy() for k, v in self.estimated_c2ws.items()}

    def set_cam_dict(self, cam_dict: dict) -> None:
        self.estimated_c2ws = {int(k): torch.from_numpy(v) for k, v in cam_dict.items()}
