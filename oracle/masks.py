"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's mask post-processing:
`mask_nms` (ovo/utils/segment_utils.py:195-259), `masks_update/filter` (:173-193) and `mask2segmap` (:12-27).
Pinned against the reference in tests/golden/masks.npz (oracle/gen_golden.py gen_masks).  Not part of the product."""
import numpy as np

f32 = np.float32


def synth_masks(M=40, H=120, W=160, seed=0):
    """Overlapping rectangles / nested boxes with scores: exercises IoU suppression, both containment rules and ties
    in the paint order."""
    rng = np.random.default_rng(seed)
    masks = np.zeros((M, H, W), bool)
    for m in range(M):
        if m % 5 == 4 and m > 0:                         # nested inside the previous mask
            ys, xs = np.nonzero(masks[m - 1])
            y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
            hh, ww = max(2, (y1 - y0) // 3), max(2, (x1 - x0) // 3)
            masks[m, y0 + hh // 2: y0 + hh // 2 + hh, x0 + ww // 2: x0 + ww // 2 + ww] = True
        elif m % 7 == 6 and m > 0:                       # near duplicate of the previous mask
            masks[m] = np.roll(masks[m - 1], 1, axis=1)
        else:
            h, w = rng.integers(8, H // 2), rng.integers(8, W // 2)
            y, x = rng.integers(0, H - h), rng.integers(0, W - w)
            masks[m, y:y + h, x:x + w] = True
    iou = rng.uniform(0.6, 1.0, M).astype(f32)
    stab = rng.uniform(0.75, 1.0, M).astype(f32)
    return masks, iou, stab


def mask_nms(masks: np.ndarray, scores: np.ndarray, iou_thr=0.7, score_thr=0.1, inner_thr=0.2) -> np.ndarray:
    """segment_utils.py:195-259.  Returns the selected ORIGINAL indices in descending-score order."""
    M = masks.shape[0]
    idx = np.argsort(-scores, kind="stable")
    s = scores[idx]
    mo = masks[idx].reshape(M, -1)
    area = mo.sum(1).astype(f32)
    inter = (mo.astype(np.int32) @ mo.astype(np.int32).T).astype(f32)            # [M,M] intersection counts
    union = area[:, None] + area[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(f32)
        ri = (inter / area[:, None]).astype(f32)          # inter / area[i]  (row i)
        rj = (inter / area[None, :]).astype(f32)          # inter / area[j]  (column j)
        inner_val = (f32(1) - (rj * ri).astype(f32)).astype(f32)
    upper = np.triu(np.ones((M, M), bool))               # pairs i <= j visited by the reference loop
    inner = np.zeros((M, M), f32)
    c1 = upper & (ri < f32(0.5)) & (rj >= f32(0.85))      # -> inner[i, j]
    inner[c1] = inner_val[c1]
    c2 = upper & (ri >= f32(0.85)) & (rj < f32(0.5))      # -> inner[j, i]
    inner.T[c2] = inner_val[c2]
    iou_u = np.triu(np.where(upper, iou, 0), 1)
    iou_max = iou_u.max(0) if M else np.zeros(0, f32)
    inner_max_u = np.triu(inner, 1).max(0)
    inner_max_l = np.tril(inner, 1).max(0)                # tril(diagonal=1): keeps the first super-diagonal too (:236)
    keep = (iou_max <= f32(iou_thr)) & (s > f32(score_thr)) & (inner_max_u <= f32(1 - inner_thr)) & (inner_max_l <= f32(1 - inner_thr))
    return idx[keep]


def masks_update(masks, iou_pred, stability, iou_thr=0.8, score_thr=0.7, inner_thr=0.5):
    """masks_update + filter (segment_utils.py:173-193): surviving ORIGINAL indices in original order."""
    sel = mask_nms(masks, (stability * iou_pred).astype(f32), iou_thr, score_thr, inner_thr)
    return np.array(sorted(sel.tolist()), np.int64)


def mask2segmap(masks: np.ndarray, stability: np.ndarray):
    """segment_utils.py:12-27: paint in descending stability, earlier masks win overlaps.
    Returns (seg_map [H,W] i32, binary_maps [M,H,W] in painted order, order)."""
    order = np.argsort(-stability, kind="stable")
    seg = -np.ones(masks.shape[1:], np.int32)
    for i, m in enumerate(order):
        seg[masks[m] & (seg == -1)] = i
    return seg, masks[order], order
