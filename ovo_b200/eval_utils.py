"""Label transfer from the predicted point map to the ground-truth mesh vertices — `match_labels_to_vtx` of the
reference (ovo/utils/eval_utils.py:13-44, called by run_eval.compute_scene_labels, run_eval.py:50) on the GPU:
the SciPy KD-tree query (k = 5) is `ovo_knn` (grid hash, exact), `torch.mode` over the five labels is `ovo_knn_mode`.
Same signature and return values as the reference; the results are returned on the device the labels came from (CPU tensors in,
CPU tensors out, as run_eval.py expects)."""
from typing import Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def knn(points: torch.Tensor, queries: torch.Tensor, k: int = 5, cell_size: float = 0.0, return_distance: bool = True):
    """Exact k nearest neighbours: points [N,3], queries [Q,3] (any float dtype / device; computed on the GPU in the
    float32 the map stores) -> (dist f64 [Q,k], idx int32 [Q,k]), ascending, like KDTree.query."""
    if not torch.cuda.is_available():
        raise RuntimeError("ovo_b200.eval_utils.knn needs a CUDA device (no CPU fallback)")
    dev = points.device if points.is_cuda else torch.device("cuda")
    p = points.to(dev, torch.float32).reshape(-1, 3).contiguous()
    q = queries.to(dev, torch.float32).reshape(-1, 3).contiguous()
    idx = torch.empty(q.shape[0], k, device=dev, dtype=torch.int32)
    dist = torch.empty(q.shape[0], k, device=dev, dtype=torch.float64) if return_distance else None
    if q.shape[0] == 0:
        return dist, idx
    with torch.cuda.device(dev):
        check(_lib.lib().ovo_knn(ptr(p), p.shape[0], ptr(q), q.shape[0], k, float(cell_size), ptr(idx), ptr(dist), stream_ptr()),
              "ovo_knn")
    return dist, idx


def knn_stats() -> dict:
    """Grid cell, cell count and number of exhaustive-fallback queries of the last `knn` call of this thread."""
    import ctypes as C
    cell, cells, fb = C.c_float(), C.c_int(), C.c_int()
    _lib.lib().ovo_knn_stats(C.byref(cell), C.byref(cells), C.byref(fb))
    return {"cell_size": round(cell.value, 5), "cells": cells.value, "fallback_queries": fb.value}


def match_labels_to_vtx(points_3d_labels: torch.Tensor, points_3d: torch.Tensor, mesh_vtx: torch.Tensor,
                        filter_unasigned: bool = True, tree: str = "kd", verbose=False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """eval_utils.py:13-44.  `tree` is accepted for compatibility (kd / ball give the same neighbours)."""
    points_3d_labels = torch.as_tensor(points_3d_labels)
    points_3d, mesh_vtx = torch.as_tensor(points_3d), torch.as_tensor(mesh_vtx)
    out_dev = points_3d_labels.device        # results go back where the labels came from (run_eval.py:50-53 feeds them to numpy)
    if filter_unasigned:
        assigned_mask = (points_3d_labels > -1).squeeze()
        if verbose:
            print(f"Assigned points {assigned_mask.sum()}, {assigned_mask.float().mean()*100:.1f}")
        points_3d_labels = points_3d_labels[assigned_mask]
        points_3d = points_3d[assigned_mask.to(points_3d.device)]
        assert len(points_3d_labels), "All points are unassigned"
    _, idx = knn(points_3d, mesh_vtx, k=5, return_distance=False)
    dev = idx.device
    labels = points_3d_labels.to(dev, torch.int32).reshape(-1).contiguous()
    mesh_labels32 = torch.empty(idx.shape[0], device=dev, dtype=torch.int32)
    check(_lib.lib().ovo_knn_mode(ptr(labels), ptr(idx), idx.shape[0], 5, ptr(mesh_labels32), stream_ptr()), "ovo_knn_mode")
    mesh_labels = mesh_labels32.to(out_dev, points_3d_labels.dtype)
    matched_instances_ids = torch.unique(mesh_labels)
    if not filter_unasigned:
        while matched_instances_ids[0] < 0:
            matched_instances_ids = matched_instances_ids[1:]
    n_instances = len(matched_instances_ids)
    instance_idxs = torch.unsqueeze(matched_instances_ids, dim=1)
    mesh_instances_masks = torch.unsqueeze(mesh_labels, dim=0).expand(n_instances, -1) == instance_idxs
    return mesh_labels, mesh_instances_masks, matched_instances_ids
