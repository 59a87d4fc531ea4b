"""Host side of the 3D association / fusion / query kernels (libovo_b200 map entry points).

`SemanticMap` replaces the body of `OVO._match_and_track_instances` + `_track_objects`
(ovo/entities/ovo.py:182-282) and carries the two map modes of SURVEY §0: the instance bank (reference
semantics) and the dense per-point bank (north-star running-mean / cosine kernels)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Frame, VoteRow, check, ptr, stream_ptr

VOTE_FIELDS = ("n_matched", "n_assigned", "n_unassigned", "mode_id", "ins_id", "is_new", "area")


class SemanticMap:
    def __init__(self, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("ovo_b200.SemanticMap needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.lib = _lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ovo_map_create(C.byref(h)), "ovo_map_create")
        self.handle = h
        self.n_slots = 64          # per-keyframe match-list slots of the handle (ovo_map::kSlots)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ovo_map_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -------------------------------------------------------------------------------------------
    def reserve(self, points: int = 0, instances: int = 0, masks: int = 0, matches: int = 0) -> None:
        """Sizes the association workspaces once (see ovo_map_reserve): no allocation pauses while the map grows."""
        with torch.cuda.device(self.device):
            check(self.lib.ovo_map_reserve(self.handle, int(points), int(instances), int(masks), int(matches)), "ovo_map_reserve")

    def depth_filter(self, depth: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """geometry_utils.depth_filter (geometry_utils.py:92-96)."""
        depth = depth.to(self.device, torch.float32).contiguous()
        if out is None:
            out = torch.empty_like(depth)
        assert out.is_contiguous() and out.shape == depth.shape and out.dtype == torch.float32
        check(self.lib.ovo_depth_filter(ptr(depth), depth.shape[0], depth.shape[1], ptr(out), stream_ptr(self.device)), "ovo_depth_filter")
        return out

    def depth_filter_batch(self, depths: torch.Tensor, out: torch.Tensor | None = None, ranges: torch.Tensor | None = None):
        """F depth maps [F,h,w] in one launch -> (filtered maps [F,h,w], raw-depth ranges [F,2]); see depth_filter / depth_range."""
        assert depths.is_cuda and depths.dtype == torch.float32 and depths.is_contiguous() and depths.dim() == 3
        if out is None:
            out = torch.empty_like(depths)
        if ranges is None:
            ranges = torch.empty(depths.shape[0], 2, device=self.device, dtype=torch.float32)
        assert out.is_contiguous() and ranges.is_contiguous() and out.shape == depths.shape
        check(self.lib.ovo_depth_filter_batch(ptr(depths), depths.shape[0], depths.shape[1], depths.shape[2], ptr(out), ptr(ranges),
                                              stream_ptr(self.device)), "ovo_depth_filter_batch")
        return out, ranges

    def depth_range(self, depth: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """[min, max] of the depth values > 0 (the raw-depth range the frustum is built from), f32 [2] on the device."""
        depth = depth.to(self.device, torch.float32).contiguous()
        if out is None:
            out = torch.empty(2, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_depth_range(ptr(depth), depth.numel(), ptr(out), stream_ptr(self.device)), "ovo_depth_range")
        return out

    def associate(self, xyz: torch.Tensor, ins_ids: torch.Tensor, depth: torch.Tensor, seg_map: torch.Tensor,
                  c2w, K, next_ins_id: int, match_th: float = 0.05, track_th: int = 100, depth_filter: bool = True,
                  rgb_depth_ratio=(), kf_slot: int = 0, n_masks: int | None = None, w2c=None):
        """One keyframe of association.  xyz [N,3] f32, ins_ids [N] i32 (updated IN PLACE), depth [h,w] f32,
        seg_map [H,W] i32, all on the device.  Returns (votes dict of np arrays [n_masks], n_matched,
        next_ins_id)."""
        assert xyz.is_cuda and ins_ids.is_cuda and depth.is_cuda and seg_map.is_cuda
        assert xyz.dtype == torch.float32 and ins_ids.dtype == torch.int32 and seg_map.dtype == torch.int32
        assert xyz.is_contiguous() and ins_ids.is_contiguous() and depth.is_contiguous() and seg_map.is_contiguous()
        if n_masks is None:
            n_masks = int(seg_map.max().item()) + 1                     # ovo.py:255
        f = self._frame(depth, seg_map, c2w, K, match_th, track_th, depth_filter, rgb_depth_ratio, n_masks, w2c)
        rows = (VoteRow * max(n_masks, 1))()
        nxt, nm = C.c_int(next_ins_id), C.c_int(0)
        check(self.lib.ovo_map_associate(self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], C.byref(f), C.byref(nxt), rows,
                                         C.byref(nm), kf_slot, stream_ptr(self.device)), "ovo_map_associate")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(-1, 8)[:n_masks]
        votes = {k: arr[:, i].copy() for i, k in enumerate(VOTE_FIELDS)}
        return votes, nm.value, nxt.value

    def associate_launch(self, xyz, ins_ids, depth, seg_map, c2w, K, next_ins_id: int, match_th=0.05, track_th=100, depth_filter=True,
                         rgb_depth_ratio=(), kf_slot: int = 0, n_masks: int | None = None, w2c=None) -> None:
        """`associate` without the host synchronisation: everything is enqueued, `associate_wait()` returns the results."""
        assert xyz.is_cuda and ins_ids.is_cuda and xyz.dtype == torch.float32 and ins_ids.dtype == torch.int32
        assert xyz.is_contiguous() and ins_ids.is_contiguous() and depth.is_contiguous() and seg_map.is_contiguous()
        if n_masks is None:
            n_masks = int(seg_map.max().item()) + 1
        f = self._frame(depth, seg_map, c2w, K, match_th, track_th, depth_filter, rgb_depth_ratio, n_masks, w2c)
        check(self.lib.ovo_map_associate_launch(self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], C.byref(f), int(next_ins_id), kf_slot,
                                                stream_ptr(self.device)), "ovo_map_associate_launch")
        self._launched = (n_masks, (xyz, ins_ids, depth, seg_map))        # keep the operands alive until the wait

    def associate_wait(self):
        n_masks, _ = self._launched
        self._launched = None
        rows = (VoteRow * max(n_masks, 1))()
        nxt, nm = C.c_int(0), C.c_int(0)
        check(self.lib.ovo_map_associate_wait(self.handle, C.byref(nxt), rows, C.byref(nm)), "ovo_map_associate_wait")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(-1, 8)[:n_masks]
        return {k: arr[:, i].copy() for i, k in enumerate(VOTE_FIELDS)}, nm.value, nxt.value

    # ---- the same association in two halves, for a map sharded over ranks (ovo_b200/sharding.py)
    def _frame(self, depth, seg_map, c2w, K, match_th, track_th, depth_filter, rgb_depth_ratio, n_masks, w2c):
        c2w = np.asarray(c2w, np.float32).reshape(4, 4)
        if w2c is None:
            w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        K = np.asarray(K, np.float32).reshape(3, 3)
        f = Frame()
        f.depth_dev = depth.data_ptr(); f.h, f.w = depth.shape
        f.seg_map_dev = seg_map.data_ptr(); f.H, f.W = seg_map.shape
        f.n_masks = n_masks
        C.memmove(f.c2w, np.ascontiguousarray(c2w).ctypes.data, 64)
        C.memmove(f.w2c, np.ascontiguousarray(w2c, np.float32).ctypes.data, 64)
        C.memmove(f.K, np.ascontiguousarray(K).ctypes.data, 36)
        f.match_th, f.track_th, f.depth_filter = float(match_th), int(track_th), int(bool(depth_filter))
        if len(rgb_depth_ratio) > 0:
            f.has_ratio, f.ratio_h, f.ratio_w, f.crop_edge = 1, float(rgb_depth_ratio[0]), float(rgb_depth_ratio[1]), int(rgb_depth_ratio[2])
        return f

    def vote(self, xyz, ins_ids, depth, seg_map, c2w, K, n_ins: int, n_masks: int, match_th=0.05, track_th=100,
             depth_filter=True, rgb_depth_ratio=(), kf_slot=0, w2c=None) -> torch.Tensor:
        """Pass 1 on this rank's points: returns the device table [n_masks*(n_ins+1) + 1] i32 (votes, n_matched)."""
        f = self._frame(depth, seg_map, c2w, K, match_th, track_th, depth_filter, rgb_depth_ratio, n_masks, w2c)
        table = torch.empty(max(n_masks, 1) * (n_ins + 1) + 1, device=self.device, dtype=torch.int32)
        self._pending = (ins_ids, n_masks)
        check(self.lib.ovo_map_vote(self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], C.byref(f), n_ins, ptr(table), kf_slot,
                                    stream_ptr(self.device)), "ovo_map_vote")
        return table

    def apply(self, table: torch.Tensor, next_ins_id: int):
        """Decisions from the (all-reduced) table + pass 2 on this rank's points -> (votes, n_matched, next_ins_id)."""
        ins_ids, n_masks = self._pending
        rows = (VoteRow * max(n_masks, 1))()
        nxt, nm = C.c_int(next_ins_id), C.c_int(0)
        check(self.lib.ovo_map_apply(self.handle, ptr(table), ptr(ins_ids), C.byref(nxt), rows, C.byref(nm), stream_ptr(self.device)), "ovo_map_apply")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(-1, 8)[:n_masks]
        return {k: arr[:, i].copy() for i, k in enumerate(VOTE_FIELDS)}, nm.value, nxt.value

    def matches(self, kf_slot: int, n_max: int) -> torch.Tensor:
        """(point index, mask index) pairs of a keyframe slot, [n,2] i32 on the device (unordered)."""
        buf = torch.empty(max(n_max, 1), 2, device=self.device, dtype=torch.int32)
        n = check(self.lib.ovo_map_get_matches(self.handle, kf_slot, ptr(buf), n_max, stream_ptr(self.device)), "ovo_map_get_matches")
        return buf[:n]

    # ---- several keyframes in one pass over the map (ovo_map_associate_batch)
    def _frames(self, depths, seg_maps, c2ws, K, n_masks, match_th, track_th, depth_filter, rgb_depth_ratio, w2cs, depth_ranges=None):
        """depth_ranges (optional, list of F device tensors f32 [2]): `depths` are already filtered (depth_filter(...)) and these
        are the raw depths' [min, max] (depth_range(...)): a sharded map gathers them instead of filtering on every rank."""
        F = len(depths)
        arr = (Frame * F)()
        for i in range(F):
            assert depths[i].is_cuda and seg_maps[i].is_cuda and depths[i].dtype == torch.float32 and seg_maps[i].dtype == torch.int32
            assert depths[i].is_contiguous() and seg_maps[i].is_contiguous()
            nm = n_masks[i] if not isinstance(n_masks, int) else n_masks
            arr[i] = self._frame(depths[i], seg_maps[i], c2ws[i], K, match_th, track_th, depth_filter, rgb_depth_ratio, int(nm),
                                 None if w2cs is None else w2cs[i])
            if depth_ranges is not None:
                arr[i].depth_range_dev = depth_ranges[i].data_ptr()
        return arr

    def associate_batch(self, xyz: torch.Tensor, ins_ids: torch.Tensor, depths, seg_maps, c2ws, K, next_ins_id: int, n_masks,
                        match_th: float = 0.05, track_th: int = 100, depth_filter: bool = True, rgb_depth_ratio=(), kf_slots=None,
                        w2cs=None, mask_ins_out: torch.Tensor | None = None, depth_ranges=None):
        """F keyframes against the map in one pass over xyz, id decisions on the device, ONE host synchronisation; the same
        results as F consecutive `associate` calls.  depths / seg_maps / c2ws: lists of F; n_masks: int or list.
        Returns (list of F votes dicts, list of F n_matched, next_ins_id); `mask_ins_out` i32 [F, stride] (optional, device)
        receives the instance id of every mask."""
        assert xyz.is_cuda and ins_ids.is_cuda and xyz.dtype == torch.float32 and ins_ids.dtype == torch.int32
        assert xyz.is_contiguous() and ins_ids.is_contiguous()
        F = len(depths)
        frames = self._frames(depths, seg_maps, c2ws, K, n_masks, match_th, track_th, depth_filter, rgb_depth_ratio, w2cs, depth_ranges)
        nms = [int(frames[i].n_masks) for i in range(F)]
        stride = max(max(nms), 1) if mask_ins_out is None else int(mask_ins_out.shape[1])
        assert mask_ins_out is None or (mask_ins_out.dtype == torch.int32 and mask_ins_out.shape[0] >= F and mask_ins_out.is_contiguous() and stride >= max(nms))
        rows = (VoteRow * (F * stride))()
        nxt, nm = C.c_int(next_ins_id), (C.c_int * F)()
        slots = None if kf_slots is None else (C.c_int * F)(*[int(x) for x in kf_slots])
        check(self.lib.ovo_map_associate_batch(self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], frames, F, slots, C.byref(nxt), rows,
                                               stride, nm, ptr(mask_ins_out), stream_ptr(self.device)), "ovo_map_associate_batch")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(F, stride, 8)
        votes = [{k: arr[f, :nms[f], i].copy() for i, k in enumerate(VOTE_FIELDS)} for f in range(F)]
        return votes, [int(x) for x in nm], nxt.value

    # the same batch in stages, for a map sharded over ranks: begin -> per keyframe (vote -> all-reduce of the table -> decide) -> end
    def batch_begin(self, xyz, ins_ids, depths, seg_maps, c2ws, K, next_ins_id: int, n_masks, tables: torch.Tensor, match_th=0.05,
                    track_th=100, depth_filter=True, rgb_depth_ratio=(), kf_slots=None, w2cs=None, depth_ranges=None):
        F = len(depths)
        frames = self._frames(depths, seg_maps, c2ws, K, n_masks, match_th, track_th, depth_filter, rgb_depth_ratio, w2cs, depth_ranges)
        assert tables.is_cuda and tables.dtype == torch.int32 and tables.is_contiguous()
        slots = None if kf_slots is None else (C.c_int * F)(*[int(x) for x in kf_slots])
        check(self.lib.ovo_map_batch_begin(self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], frames, F, slots, int(next_ins_id),
                                           ptr(tables), tables.numel(), stream_ptr(self.device)), "ovo_map_batch_begin")
        self._batch = (ins_ids, [int(frames[i].n_masks) for i in range(F)], tables)

    def batch_vote(self, f: int) -> torch.Tensor:
        """Votes of keyframe f on this rank's points -> the view of `tables` that holds its table (sum it over the ranks in place)."""
        ins_ids, _, tables = self._batch
        p, n = C.c_void_p(), C.c_int(0)
        check(self.lib.ovo_map_batch_vote(self.handle, f, ptr(ins_ids), C.byref(p), C.byref(n), stream_ptr(self.device)), "ovo_map_batch_vote")
        off = (p.value - tables.data_ptr()) // 4
        return tables[off: off + n.value]

    def batch_decide(self, f: int) -> None:
        check(self.lib.ovo_map_batch_decide(self.handle, f, stream_ptr(self.device)), "ovo_map_batch_decide")

    def batch_end(self, mask_ins_out: torch.Tensor | None = None):
        ins_ids, nms, _ = self._batch
        F = len(nms)
        stride = max(max(nms), 1) if mask_ins_out is None else int(mask_ins_out.shape[1])
        rows = (VoteRow * (F * stride))()
        nxt, nm = C.c_int(0), (C.c_int * F)()
        check(self.lib.ovo_map_batch_end(self.handle, ptr(ins_ids), C.byref(nxt), rows, stride, nm, ptr(mask_ins_out),
                                         stream_ptr(self.device)), "ovo_map_batch_end")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(F, stride, 8)
        votes = [{k: arr[f, :nms[f], i].copy() for i, k in enumerate(VOTE_FIELDS)} for f in range(F)]
        return votes, [int(x) for x in nm], nxt.value

    @staticmethod
    def batch_tables_size(next_ins_id: int, n_masks) -> int:
        """ints `tables` needs for a batch whose keyframes have n_masks[f] masks (ovo_map_batch_begin)."""
        n, bound = 0, int(next_ins_id)
        for nm in n_masks:
            n += (4 + max(int(nm), 1) * (bound + 1) + 3) // 4 * 4
            bound += int(nm)
        return n

    def fuse_dense(self, kf_slot: int, bank: torch.Tensor, bank_lo: torch.Tensor, counts: torch.Tensor, feats: torch.Tensor,
                   mask_row: torch.Tensor):
        """Per-point running mean in the two-plane bank: bank / bank_lo [N,D] bf16 (mean = bank + bank_lo; `bank` alone is the
        query operand), counts [N] i32, feats [R,D] f32, mask_row [n_masks] i32."""
        assert bank.dtype == torch.bfloat16 and bank_lo.dtype == torch.bfloat16 and counts.dtype == torch.int32 and feats.dtype == torch.float32
        assert mask_row.dtype == torch.int32 and bank.is_contiguous() and bank_lo.is_contiguous() and feats.is_contiguous()
        assert bank_lo.shape == bank.shape and feats.shape[1] == bank.shape[1]
        check(self.lib.ovo_map_fuse_dense(self.handle, kf_slot, ptr(bank), ptr(bank_lo), ptr(counts), bank.shape[0], bank.shape[1],
                                          ptr(feats), feats.shape[0], ptr(mask_row), mask_row.shape[0], stream_ptr(self.device)),
              "ovo_map_fuse_dense")

    def fuse_dense_batch(self, kf_slots, bank: torch.Tensor, bank_lo: torch.Tensor, counts: torch.Tensor, feats: torch.Tensor,
                         mask_row: torch.Tensor):
        """Several keyframes in one pass over the bank (a point's descriptors are summed in f32, then ONE mean update).
        feats [R,D] f32 = descriptors of all keyframes, mask_row [len(kf_slots), n_masks] i32 -> row of feats or -1."""
        assert bank.dtype == torch.bfloat16 and bank_lo.dtype == torch.bfloat16 and counts.dtype == torch.int32 and feats.dtype == torch.float32
        assert mask_row.dtype == torch.int32 and mask_row.dim() == 2 and mask_row.shape[0] == len(kf_slots) and mask_row.is_contiguous()
        arr = (C.c_int * len(kf_slots))(*[int(s) for s in kf_slots])
        assert feats.is_contiguous() and feats.shape[1] == bank.shape[1] and bank_lo.shape == bank.shape and bank.is_contiguous() and bank_lo.is_contiguous()
        check(self.lib.ovo_map_fuse_dense_batch(self.handle, arr, len(kf_slots), ptr(bank), ptr(bank_lo), ptr(counts), bank.shape[0],
                                                bank.shape[1], ptr(feats), feats.shape[0], ptr(mask_row), mask_row.shape[1],
                                                stream_ptr(self.device)), "ovo_map_fuse_dense_batch")

    def query_dense(self, bank: torch.Tensor, text: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """clip_cosine_similarity over the dense bank: [N,D] bf16 x [Q,D] f32 -> [N,Q] f32."""
        assert bank.dtype == torch.bfloat16 and bank.is_contiguous()
        text = text.to(self.device, torch.float32).contiguous()
        N, D = bank.shape
        Q = text.shape[0]
        if out is None:
            out = torch.empty(N, Q, device=self.device, dtype=torch.float32)
        check(self.lib.ovo_query_dense(self.handle, ptr(bank), N, D, ptr(text), Q, ptr(out), stream_ptr(self.device)), "ovo_query_dense")
        return out

    def query_instances(self, bank: torch.Tensor, text: torch.Tensor, rows: torch.Tensor | None = None) -> torch.Tensor:
        """clip_cosine_similarity over the instance bank in f32: bank[rows] [I,D] x [Q,D] -> [I,Q]."""
        bank = bank.to(self.device, torch.float32).contiguous()
        text = text.to(self.device, torch.float32).contiguous()
        n = bank.shape[0] if rows is None else rows.shape[0]
        out = torch.empty(n, text.shape[0], device=self.device, dtype=torch.float32)
        check(self.lib.ovo_query_instances(ptr(bank), ptr(rows), n, bank.shape[1], ptr(text), text.shape[0], ptr(out),
                                           stream_ptr(self.device)), "ovo_query_instances")
        return out

    def merge_masks(self, masks: torch.Tensor, group: torch.Tensor, n_out: int):
        """_fuse_masks_with_same_ins_id: masks uint8 [M,H,W], group i32 [M] -> (out uint8 [R,H,W], areas i32 [R])."""
        assert masks.dtype == torch.uint8 and masks.is_contiguous() and group.dtype == torch.int32
        M, H, W = masks.shape
        out = torch.empty(n_out, H, W, device=self.device, dtype=torch.uint8)
        areas = torch.empty(n_out, device=self.device, dtype=torch.int32)
        check(self.lib.ovo_merge_masks(ptr(masks), M, H, W, ptr(group), n_out, ptr(out), ptr(areas), stream_ptr(self.device)), "ovo_merge_masks")
        return out, areas

    def fuse_views(self, store: torch.Tensor, idx: torch.Tensor, off: torch.Tensor, mode: int, bank: torch.Tensor,
                   out_rows: torch.Tensor, chosen: torch.Tensor | None = None):
        """Instance3D.update_clip batched: see ovo_fuse_views in include/ovo_b200.h."""
        assert store.dtype == torch.float32 and bank.dtype == torch.float32 and idx.dtype == torch.int32
        check(self.lib.ovo_fuse_views(ptr(store), store.shape[1], ptr(idx), ptr(off), out_rows.shape[0], mode, ptr(bank),
                                      ptr(out_rows), ptr(chosen), stream_ptr(self.device)), "ovo_fuse_views")

    def mask_nms(self, masks: torch.Tensor, scores: torch.Tensor, iou_thr=0.8, score_thr=0.7, inner_thr=0.5) -> torch.Tensor:
        """masks_update/mask_nms/filter (segment_utils.py:173-259): masks [M,H,W] bool/uint8, scores [M] f32 ->
        keep [M] bool in the original order."""
        m8 = masks.to(self.device).to(torch.uint8).contiguous()
        sc = scores.to(self.device, torch.float32).contiguous()
        keep = torch.empty(m8.shape[0], device=self.device, dtype=torch.uint8)
        check(self.lib.ovo_mask_nms(ptr(m8), ptr(sc), m8.shape[0], m8.shape[1], m8.shape[2], float(iou_thr), float(score_thr),
                                    float(inner_thr), ptr(keep), stream_ptr(self.device)), "ovo_mask_nms")
        return keep.bool()

    def mask2segmap(self, masks: torch.Tensor, stability: torch.Tensor):
        """mask2segmap (segment_utils.py:12-27) -> (seg_map [H,W] i32, binary_maps [M,H,W] bool in painted order, order)."""
        m8 = masks.to(self.device).to(torch.uint8).contiguous()
        st = stability.to(self.device, torch.float32).contiguous()
        M, H, W = m8.shape
        seg = torch.empty(H, W, device=self.device, dtype=torch.int32)
        maps = torch.empty_like(m8)
        order = torch.empty(M, device=self.device, dtype=torch.int32)
        check(self.lib.ovo_mask2segmap(ptr(m8), ptr(st), M, H, W, ptr(seg), ptr(maps), ptr(order), stream_ptr(self.device)), "ovo_mask2segmap")
        return seg, maps.bool(), order

    def classify(self, sim: torch.Tensor, th: float = 0.0):
        """OVO.classify_instances' argmax + threshold (ovo.py:486-491)."""
        sim = sim.contiguous()
        cls = torch.empty(sim.shape[0], device=self.device, dtype=torch.int32)
        conf = torch.empty(sim.shape[0], device=self.device, dtype=torch.float32)
        check(self.lib.ovo_classify(ptr(sim), sim.shape[0], sim.shape[1], float(th), ptr(cls), ptr(conf), stream_ptr(self.device)), "ovo_classify")
        return cls, conf

    def bank_add_views(self, bank: torch.Tensor, store: torch.Tensor, quads: torch.Tensor, idx: torch.Tensor):
        """bank[b] = (n*bank[b] + sum store[idx[i0:i1]]) / (n + i1 - i0) for int32 quads [m,4] = (b, n, i0, i1): a few more
        views per instance (avg_pooling) without re-reading the views already fused."""
        assert bank.is_cuda and store.is_cuda and quads.is_cuda and idx.is_cuda
        assert quads.dtype == torch.int32 and idx.dtype == torch.int32 and quads.is_contiguous() and idx.is_contiguous()
        check(self.lib.ovo_bank_add_views(ptr(bank), bank.shape[1], ptr(store), ptr(quads), ptr(idx), quads.shape[0], stream_ptr(self.device)),
              "ovo_bank_add_views")

    def bank_update_mean(self, bank: torch.Tensor, counts: torch.Tensor, feats: torch.Tensor, rows: torch.Tensor):
        """avg_pooling running mean on the instance bank (instance3d.py:19-21)."""
        assert bank.dtype == torch.float32 and counts.dtype == torch.int32 and rows.dtype == torch.int32
        feats = feats.to(self.device, torch.float32).contiguous()
        check(self.lib.ovo_bank_update_mean(ptr(bank), ptr(counts), bank.shape[1], ptr(feats), ptr(rows), rows.shape[0],
                                            stream_ptr(self.device)), "ovo_bank_update_mean")
