"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32, fixed operation order) of the reference's
3D association + fusion path.  Not part of the product; see oracle/encoder.py header for the rules.

  gu  = ovo/utils/geometry_utils.py      ovo = ovo/entities/ovo.py      i3d = ovo/entities/instance3d.py

Floating point feeds integer decisions here (pixel rounding, plane tests, depth threshold).  The
reference evaluates them with torch.einsum / matmul / conv2d whose accumulation order is BLAS-defined;
this restatement FIXES one order (left-to-right, every product and sum rounded to float32, no FMA) and
the CUDA kernels use the same order with __fmul_rn/__fadd_rn, so CUDA == oracle bit-for-bit, and
oracle == reference is pinned on the golden fixtures (tests/golden/assoc_*.npz, generated from the
reference with oracle/gen_golden.py; mismatches counted there).
"""
import numpy as np

f32 = np.float32


def _dot4(m_row, x, y, z, w=None):
    """((m0*x + m1*y) + m2*z) + m3*w  in float32, left to right."""
    acc = f32(m_row[0]) * x
    acc = acc + f32(m_row[1]) * y
    acc = acc + f32(m_row[2]) * z
    if w is None:
        acc = acc + f32(m_row[3])
    else:
        acc = acc + f32(m_row[3]) * w
    return acc.astype(f32) if isinstance(acc, np.ndarray) else f32(acc)


def frustum_corners(depth: np.ndarray, c2w: np.ndarray, K: np.ndarray) -> np.ndarray:
    """gu:99-129.  depth [h,w] f32, c2w [4,4] f32, K [3,3] f32 -> corners [8,3] f32."""
    h, w = depth.shape
    valid = depth[depth > 0]
    dmin, dmax = f32(valid.min()), f32(valid.max())
    px = np.array([0, w, 0, w, 0, w, 0, w], f32)
    py = np.array([0, 0, h, h, 0, 0, h, h], f32)
    pz = np.array([dmin] * 4 + [dmax] * 4, f32)
    x = ((px - f32(K[0, 2])) * pz / f32(K[0, 0])).astype(f32)
    y = ((py - f32(K[1, 2])) * pz / f32(K[1, 1])).astype(f32)
    out = np.zeros((8, 3), f32)
    for r in range(3):
        out[:, r] = _dot4(c2w[r].astype(f32), x, y, pz)
    return out


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], f32)


def frustum_planes(c: np.ndarray) -> np.ndarray:
    """gu:163-207.  corners [8,3] -> planes [6,4] (near, far, left, right, top, bottom), inside iff n.p+d<=0."""
    pairs = [(2, 0, 1, 0), (6, 4, 5, 4), (4, 0, 2, 0), (7, 3, 1, 3), (5, 1, 3, 1), (6, 2, 0, 2)]
    planes = np.zeros((6, 4), f32)
    for i, (a, b, cc, d) in enumerate(pairs):
        n = _cross((c[a] - c[b]).astype(f32), (c[cc] - c[d]).astype(f32))
        dd = -f32(f32(f32(n[0] * c[i][0]) + f32(n[1] * c[i][1])) + f32(n[2] * c[i][2]))
        planes[i, :3], planes[i, 3] = n, dd
    return planes


def frustum_mask(xyz: np.ndarray, corners: np.ndarray, planes: np.ndarray) -> np.ndarray:
    """gu:252-277: AABB broad phase then 6 half-space tests.  xyz [N,3] f32 -> bool [N]."""
    lo, hi = corners.min(0), corners.max(0)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    m = (x >= lo[0]) & (x <= hi[0]) & (y >= lo[1]) & (y <= hi[1]) & (z >= lo[2]) & (z <= hi[2])
    for p in planes:
        m &= _dot4(p, x, y, z) <= 0
    return m


# torchvision _get_gaussian_kernel1d(7, 2.5, float32) bit patterns (torch.exp differs from np.exp by 1-2 ulp,
# so the reference's values are pinned here and in ovo_b200/csrc/map.cu; checked in tests/test_oracle_pin.py)
_K1D_7_25_BITS = np.array([1035802123, 1041042090, 1043549328, 1044527997, 1043549328, 1041042090, 1035802123],
                          np.uint32)


def gaussian_kernel1d(k: int = 7, sigma: float = 2.5) -> np.ndarray:
    """torchvision _get_gaussian_kernel1d."""
    if k == 7 and sigma == 2.5:
        return _K1D_7_25_BITS.view(f32).copy()
    half = (k - 1) * 0.5
    x = np.linspace(-half, half, k, dtype=f32)
    pdf = np.exp(-0.5 * (x / f32(sigma)) ** 2).astype(f32)
    return (pdf / pdf.sum()).astype(f32)


def depth_filter(depth: np.ndarray, k: int = 7, sigma: float = 2.5, th: float = 0.05) -> np.ndarray:
    """gu:92-96: 7x7 gaussian blur (reflect padding, 2-D kernel = outer(k1d,k1d)), |d-blur|>th -> -1.
    Accumulation order fixed: rows top to bottom, columns left to right."""
    k1 = gaussian_kernel1d(k, sigma)
    k2 = np.outer(k1, k1).astype(f32)
    r = k // 2
    pad = np.pad(depth.astype(f32), r, mode="reflect")
    h, w = depth.shape
    acc = np.zeros((h, w), f32)
    for dy in range(k):
        for dx in range(k):
            acc = (acc + k2[dy, dx] * pad[dy:dy + h, dx:dx + w]).astype(f32)
    hf = np.abs(depth.astype(f32) - acc)
    return np.where(hf > f32(th), f32(-1), depth.astype(f32)).astype(f32)


def project_match(xyz: np.ndarray, depth: np.ndarray, w2c: np.ndarray, K: np.ndarray, th: float):
    """gu:26-89 on an already culled subset.  Returns (ok bool[n], u int32[n], v int32[n])."""
    h, w = depth.shape
    x, y, z = xyz[:, 0].astype(f32), xyz[:, 1].astype(f32), xyz[:, 2].astype(f32)
    w2c = w2c.astype(f32)
    lx, ly, lz, lw = (_dot4(w2c[r], x, y, z) for r in range(4))
    with np.errstate(divide="ignore", invalid="ignore"):
        X, Y, Z = (lx / lw).astype(f32), (ly / lw).astype(f32), (lz / lw).astype(f32)
        K = K.astype(f32)
        ph = [((f32(K[r, 0]) * X + f32(K[r, 1]) * Y).astype(f32) + f32(K[r, 2]) * Z).astype(f32) for r in range(3)]
        uf, vf = np.rint((ph[0] / ph[2]).astype(f32)), np.rint((ph[1] / ph[2]).astype(f32))
    finite = np.isfinite(uf) & np.isfinite(vf) & (np.abs(uf) < 2e9) & (np.abs(vf) < 2e9)
    u = np.where(finite, uf, -1).astype(np.int64).astype(np.int32)
    v = np.where(finite, vf, -1).astype(np.int64).astype(np.int32)
    inpl = finite & (u < w) & (v < h) & (u >= 0) & (v >= 0)
    d = depth[np.clip(v, 0, h - 1), np.clip(u, 0, w - 1)].astype(f32)
    ok = inpl & (np.abs(lz - d) < f32(th)) & (d != 0)
    return ok, u, v


def associate(xyz, ins_ids, depth, seg_map, c2w, w2c, K, match_th, use_depth_filter=True, rgb_depth_ratio=()):
    """ovo:204-224: cull -> (depth filter) -> project/match -> seg lookup.
    Returns seg_of_pt int32[N] (-2 = not matched, else seg_map value which may be -1)."""
    N = xyz.shape[0]
    seg_of_pt = np.full(N, -2, np.int32)
    corners = frustum_corners(depth, c2w, K)
    planes = frustum_planes(corners)
    fm = frustum_mask(xyz, corners, planes)
    d = depth_filter(depth) if use_depth_filter else depth
    idx = np.nonzero(fm)[0]
    ok, u, v = project_match(xyz[idx], d, w2c, K, match_th)
    u, v = u[ok], v[ok]
    if len(rgb_depth_ratio) > 0:                      # ovo:218-221
        u = u + int(rgb_depth_ratio[-1]); v = v + int(rgb_depth_ratio[-1])
        v = (v.astype(f32) * f32(rgb_depth_ratio[0])).astype(np.int32)   # int tensor * python float -> f32 in torch
        u = (u.astype(f32) * f32(rgb_depth_ratio[1])).astype(np.int32)
    seg_of_pt[idx[ok]] = seg_map[v, u]
    return seg_of_pt, fm


def track(ins_ids: np.ndarray, seg_of_pt: np.ndarray, seg_map: np.ndarray, track_th: int, next_ins_id: int):
    """ovo:240-282 without the Instance3D side effects.
    Returns (ins_ids_new int32[N], rows list of dict per mask, next_ins_id).
    row: mask, n_matched, n_assigned, n_unassigned, mode_id, ins_id (-1 none), is_new, area."""
    ins = ins_ids.copy()
    n_masks = int(seg_map.max()) + 1
    rows = []
    for m in range(n_masks):
        pts = np.nonzero(seg_of_pt == m)[0]
        assigned = ins_ids[pts] > -1
        row = dict(mask=m, n_matched=len(pts), n_assigned=int(assigned.sum()), n_unassigned=int((~assigned).sum()),
                   mode_id=-1, ins_id=-1, is_new=0, area=int((seg_map == m).sum()))
        if row["n_assigned"] > 0:
            vals, cnt = np.unique(ins_ids[pts[assigned]], return_counts=True)
            row["mode_id"] = int(vals[np.argmax(cnt)])           # ties -> smallest id (torch.mode on CPU)
        if len(pts) > track_th:                                   # ovo:258
            if row["n_assigned"] > track_th:                      # ovo:263-269
                row["ins_id"] = row["mode_id"]
            elif row["n_unassigned"] > track_th:                  # ovo:271-276
                row["ins_id"] = next_ins_id; row["is_new"] = 1; next_ins_id += 1
            if row["ins_id"] > -1:                                # ovo:278-280
                ins[pts[~assigned]] = row["ins_id"]
        rows.append(row)
    return ins, rows, next_ins_id


def fuse_masks(binary_maps: np.ndarray, rows: list):
    """ovo:284-324 with n_top_views<=0 or every keyframe in the top-k (k_top_views=10000 default):
    masks voted to the same instance are OR-ed into the first; order = first-vote order.
    Returns (matched_ins_ids list, fused maps [M',H,W], mask_row int32[n_masks] (-1 = dropped))."""
    order, groups = [], {}
    for r in rows:
        if r["ins_id"] > -1:
            if r["ins_id"] not in groups:
                groups[r["ins_id"]] = []; order.append(r["ins_id"])
            groups[r["ins_id"]].append(r["mask"])
    bm = binary_maps.copy()
    mask_row = np.full(len(rows), -1, np.int32)
    idxs = []
    for j, ins in enumerate(order):
        first = groups[ins][0]
        for other in groups[ins][1:]:
            bm[first] |= bm[other]
        for mm in groups[ins]:
            mask_row[mm] = j
        idxs.append(first)
    return order, bm[idxs] if len(idxs) else bm[:0], mask_row


def avg_pooling(clips: np.ndarray) -> np.ndarray:
    """i3d:19-21: mean over keyframe descriptors, NOT re-normalised."""
    return clips.astype(f32).mean(axis=0, dtype=f32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """f32 -> nearest bf16 (ties to even), returned as f32 (what __floats2bfloat162_rn / torch .bfloat16() do)."""
    u = np.ascontiguousarray(x, dtype=f32).view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).astype(np.uint32)
    return r.view(f32)


def dense_update(hi: np.ndarray, lo: np.ndarray, count: int, descs: np.ndarray):
    """Dense per-point analogue of i3d:19-21 (north-star F6) on the two-plane bank: `hi` [D] = the running mean rounded to bf16
    (the query operand), `lo` [D] = bf16(mean - hi).  One update by the k descriptors a pass brings for the point
    (`descs` [k, D] f32, keyframe order), every operation rounded to f32:
        f = hi + lo;  s = sum of the bf16-rounded descriptors in order;  c' = c + k;  f' = f + (s - k*f) * (1/c')
        hi' = bf16(f');  lo' = bf16(f' - hi')
    With k = 1 this is the running mean f += (e - f)/c' of one keyframe.  Returns (hi', lo', c')."""
    k = len(descs)
    f = (hi.astype(f32) + lo.astype(f32)).astype(f32)
    e = bf16_round(descs)
    s = e[0].copy()
    for i in range(1, k):
        s = (s + e[i]).astype(f32)
    c1 = count + k
    inv = f32(1) / f32(c1)
    t = (f32(k) * f).astype(f32)
    f1 = (f + ((s - t).astype(f32) * inv).astype(f32)).astype(f32)
    hi1 = bf16_round(f1)
    lo1 = bf16_round((f1 - hi1).astype(f32))
    return hi1, lo1, c1


def dense_fuse(hi: np.ndarray, lo: np.ndarray, counts: np.ndarray, seg_of_pt_per_kf, mask_row_per_kf, feats: np.ndarray):
    """One pass of ovo_map_fuse_dense(_batch): keyframes in order, `seg_of_pt_per_kf[f]` [N] = mask of each point (-1 none),
    `mask_row_per_kf[f]` [n_masks] = row of `feats` or -1.  Updates hi / lo / counts in place."""
    N = hi.shape[0]
    rows = np.full((len(seg_of_pt_per_kf), N), -1, np.int64)
    for f, (seg, mr) in enumerate(zip(seg_of_pt_per_kf, mask_row_per_kf)):
        ok = seg >= 0
        rows[f, ok] = np.asarray(mr)[seg[ok]]
    for p in np.nonzero((rows >= 0).any(axis=0))[0]:
        r = rows[:, p]
        hi[p], lo[p], counts[p] = dense_update(hi[p], lo[p], int(counts[p]), feats[r[r >= 0]])
    return hi, lo, counts


# ---------------------------------------------------------------------------------------------------
# The same vote (ovo:240-282) decomposed for a map sharded over several ranks: every rank builds
# the vote table of its own points, the tables are summed, and every rank takes the same decisions.
# ---------------------------------------------------------------------------------------------------
def vote_table(ins_ids: np.ndarray, seg_of_pt: np.ndarray, n_masks: int, n_ins: int) -> np.ndarray:
    """[n_masks, n_ins+1] int32: column 0 = matched points without an instance, column 1+id = points carrying id."""
    t = np.zeros((n_masks, n_ins + 1), np.int32)
    sel = seg_of_pt >= 0
    np.add.at(t, (seg_of_pt[sel], np.where(ins_ids[sel] >= 0, ins_ids[sel] + 1, 0)), 1)
    return t


def decide_from_table(table: np.ndarray, areas: np.ndarray, track_th: int, next_ins_id: int):
    """rows (same dicts as track()) and the new next_ins_id from a (summed) vote table."""
    rows = []
    for m in range(table.shape[0]):
        n_un, assigned = int(table[m, 0]), table[m, 1:]
        n_as = int(assigned.sum())
        row = dict(mask=m, n_matched=n_un + n_as, n_assigned=n_as, n_unassigned=n_un, mode_id=-1, ins_id=-1, is_new=0,
                   area=int(areas[m]))
        if n_as > 0:
            row["mode_id"] = int(np.argmax(assigned))            # first maximum = smallest id on ties
        if row["n_matched"] > track_th:
            if n_as > track_th:
                row["ins_id"] = row["mode_id"]
            elif n_un > track_th:
                row["ins_id"] = next_ins_id; row["is_new"] = 1; next_ins_id += 1
        rows.append(row)
    return rows, next_ins_id


def apply_decisions(ins_ids: np.ndarray, seg_of_pt: np.ndarray, rows: list) -> np.ndarray:
    out = ins_ids.copy()
    mask_ins = np.array([r["ins_id"] for r in rows] + [-1], np.int32)
    sel = (seg_of_pt >= 0) & (ins_ids == -1)
    new = mask_ins[seg_of_pt[sel]]
    out[sel] = np.where(new >= 0, new, -1)
    return out
