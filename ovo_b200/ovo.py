"""`OVO` — the semantic module of the reference (ovo/entities/ovo.py:14-575) re-built on the sm_100a kernels,
with the same constructor, methods, attributes and return conventions so `ovo/entities/ovomapping.py` and
`ovo/slam/*` can call it unchanged.

What is different underneath:
  * association + instance vote run as two streaming kernels with ONE host sync per keyframe
    (reference: ~15 ATen kernels + a Python loop with `.item()`/`.tolist()` per mask, ovo.py:255-280);
  * per-keyframe descriptors and the fused instance descriptors live in device-resident tables; all
    instances touched by a keyframe are fused in one launch (reference: CPU tensors in Python dicts);
  * optional dense per-point map (`semantic.dense_map: True`): every matched point keeps a running mean of
    the region features it fell into, queried with the tcgen05 cosine kernel (north-star F6/Q2).
Reference behaviours that look odd are kept on purpose: `to_update` gating, `[1,D]` vs
`[D]` descriptor shapes, non-renormalised averages, `query` returning [n_obj, n_query].
"""
import pprint
import time
from collections import deque
from typing import Any, Dict, List, Tuple

import numpy as np
import torch

from .clip_generator import CLIPGenerator
from .instance3d import Instance3D, FUSION_MODES
from .map import SemanticMap
from .mask_generator import MaskGenerator


class OVO:
    def __init__(self, config: Dict[str, Any], logger, scene_name: str | None = None,
                 cam_intrinsics: torch.Tensor | None = None, eval: bool = False, device="cuda",
                 clip_generator: CLIPGenerator | None = None) -> None:
        if not eval:
            assert cam_intrinsics is not None, "Camera intrinsics required for reconstruction!"
        if not torch.cuda.is_available():
            raise RuntimeError("ovo_b200.OVO needs a CUDA device (there is no CPU fallback)")
        config["sam"]["multi_crop"] = False if config["clip"]["embed_type"] == "vanilla" else True
        self.cam_intrinsics = cam_intrinsics
        self.config = config
        self.logger = logger
        self.debug_info = config.get("debug_info", False)
        self.device = device
        self.n_top_views = config["clip"].get("k_top_views", 0)
        Instance3D.n_top_kf = self.n_top_views
        Instance3D.set_fusion(config["clip"].get("fusion", "l1_medoid"), config["clip"].get("mv_fuser_ckpt"))
        if "mask_res" in config["sam"] and "mask_res" not in config["clip"]:
            config["clip"]["mask_res"] = config["sam"]["mask_res"]

        self.clip_generator = clip_generator or CLIPGenerator(config["clip"], device=device)
        self.mask_generator = None if eval else MaskGenerator(config["sam"], scene_name, device=device)
        self.keyframes = {"ins_descriptors": dict(), "frame_id": list(), "ins_maps": list()}
        self._desc_epoch = 0      # bumped whenever stored descriptors are dropped / rewritten: incremental fusion falls back to a full one
        self.keyframes_queue = deque([])
        self.objects: Dict[int, Instance3D] = dict()
        self._time_cache = []
        self.next_ins_id = 0
        self.kf_id = 0
        self.th_centroid = config.get("th_centroid", 1.5)
        self.th_cossim = config.get("th_cossim", 0.81)
        self.th_points = config.get("th_points", 0.1)

        # device-resident tables
        self._dev = self.clip_generator.encoder.device
        self.semmap = SemanticMap(self._dev)
        if config.get("reserve_points"):      # optional, new: a stream whose map keeps growing sizes the workspaces up front
            self.semmap.reserve(points=int(config["reserve_points"]), instances=int(config.get("bank_capacity", 4096)),
                                masks=int(config.get("reserve_masks", 256)), matches=int(config.get("reserve_matches", 0)))
        # Descriptor work (ViT, pooling, fusion) is enqueued on its own stream: association synchronises the host once
        # per keyframe, and that wait must not include the encoder of earlier keyframes.  Readers of descriptors
        # (query, get_objs_clips, capture_dict, ...) first make the current stream wait for this one.
        self._enc_stream = torch.cuda.Stream(device=self._dev)
        D = self.clip_generator.clip_dim
        # `store_capacity` / `bank_capacity` (optional, new): initial rows; both tables double when full (a pause for the allocation,
        # the copy and re-pointing the instances' descriptor views: up to ~100 ms when keyframes are processed one at a time,
        # bench.py e2e.batch_keyframes_1), so a latency-sensitive stream reserves them up front.  32768 rows = 128 MB = 680
        # keyframes of 48 masks before the first doubling
        self._store = torch.zeros(int(config.get("store_capacity", 32768)), D, device=self._dev, dtype=torch.float32)   # per-keyframe descriptors
        self._store_n = 0
        self._bank = torch.zeros(int(config.get("bank_capacity", 4096)), D, device=self._dev, dtype=torch.float32)    # fused instance descriptors
        self._bank_n = 0
        self._rows_cache = None
        # map sharded over the ranks of a torch.distributed group (new, `semantic.shard_map: True`): `map_data` is then
        # THIS rank's shard (points whose voxel hashes to it, ovo_b200.sharding.shard_of_points), every rank makes the same calls
        # with the same frames; per keyframe the vote tables are summed over the ranks (one all-reduce), every rank takes the same
        # decisions, keyframe k's descriptors are computed by rank k % world and broadcast, the instance registry / bank are
        # replicated, the dense bank is sharded like the map
        self.sharded = bool(config.get("shard_map", False))
        self._group = config.get("shard_group", None)
        self._rank, self._world = 0, 1
        if self.sharded:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("ovo_b200: semantic.shard_map needs an initialised torch.distributed process group (one process per GPU)")
            self._rank, self._world = dist.get_rank(self._group), dist.get_world_size(self._group)
        # dense per-point mode
        self.dense = bool(config.get("dense_map", False))
        self._dense_bank = None        # [N, D] bf16: the running mean rounded to bf16 (the operand of the dense query)
        self._dense_bank_lo = None     # [N, D] bf16: mean - _dense_bank (the bits the first plane cannot hold)
        self._dense_counts = None
        if config.get("gc_freeze", False):
            # optional, new: a generation-2 pass of Python's cyclic GC walks every tracked object of the process (~10^6 after importing
            # torch): an isolated frame of 40+ ms in a stream that otherwise takes 4 ms per frame (measured, bench.py `stream`).
            # Freezing the start-up heap takes it out of the collector's reach; objects created from here on are collected as usual.
            import gc
            gc.collect()
            gc.freeze()
        if config.get("verbose", True):
            print('Semantic config')
            pprint.PrettyPrinter().pprint(config)

    # ------------------------------------------------------------------------------------------ device moves
    def to(self, device: str) -> None:
        return self.cuda() if "cuda" in device else self.cpu()

    def cpu(self) -> None:
        self.device = "cpu"
        self.clip_generator.cpu()
        if self.mask_generator is not None:
            self.mask_generator.cpu()

    def cuda(self) -> None:
        self.device = "cuda"
        self.clip_generator.cuda()
        if self.mask_generator is not None:
            self.mask_generator.cuda()

    def profil(func):
        """Same timing hook as the reference (ovo.py:101-119)."""
        def wrapper(self, *args, **kwargs):
            if self.config.get("log", False):
                torch.cuda.synchronize()
                t0 = time.time()
                out = func(self, *args, **kwargs)
                torch.cuda.synchronize()
                self._time_cache.append(time.time() - t0)
                return out
            return func(self, *args, **kwargs)
        return wrapper

    # ------------------------------------------------------------------------------------------ keyframe: association
    def detect_and_track_objects(self, frame_data, map_data, c2w: torch.Tensor) -> torch.Tensor:
        """ovo.py:121-166.  frame_data = (frame_id, image HxWx3 u8, depth hxw f32, rgb_depth_ratio tuple);
        map_data = (xyz [N,3] f32, ids [N(,1)] i32, obj_ids [N] i32) on the device.  Returns the updated
        per-point instance ids [N] i32, or None when no mask was produced."""
        frame_id, image = frame_data[:2]
        seg_maps, binary_maps = self._get_masks(image, frame_id)
        if len(seg_maps) == 0:
            print(f"No mask segmented in {frame_id}!")
            return None
        if self.dense and len(self.keyframes_queue) >= self.semmap.n_slots - 1:
            # the match list of a queued keyframe lives in one of the map handle's slots (kf_id % n_slots) until its dense fusion ran
            raise RuntimeError(f"ovo_b200: {len(self.keyframes_queue)} keyframes are queued for descriptors; dense_map keeps at most "
                               f"{self.semmap.n_slots - 1} (lower kf_queue_delay or call compute_semantic_info more often)")
        last_id = self.next_ins_id
        matched_ins_ids, binary_maps, n_matched_points, updated, extra = self._match_and_track_instances(
            frame_data[1:], map_data, c2w, seg_maps, binary_maps)
        self.keyframes_queue.append([matched_ins_ids, binary_maps, image, self.kf_id, extra])
        self.kf_id += 1
        if self.config.get("log", False):
            self.keyframes["frame_id"].append(frame_id)
            self.logger.log_ovo_stats({"frame_id": frame_id, "n_obj": [self.next_ins_id - last_id],
                                       "n_matches": n_matched_points, "t_sam": round(self._time_cache[0], 2),
                                       "t_obj": round(self._time_cache[1], 3)}, print_output=True)
            self._time_cache = []
        return updated

    @profil
    def _get_masks(self, image: np.ndarray, frame_id: int):
        return self.mask_generator.get_masks(image, frame_id)

    @profil
    def _match_and_track_instances(self, frame_data, map_data, c2w, seg_map, binary_maps):
        """ovo.py:182-238 on the device: one fused cull+project+match+vote pass, a second pass that gives
        unassigned points their instance, then the host bookkeeping the reference does per mask."""
        kf_id = self.kf_id
        image, depth, rgb_depth_ratio = frame_data
        points_3d, points_ids, points_ins_ids = map_data
        dev = self._dev
        depth_d = torch.as_tensor(depth, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
        seg_d = seg_map.to(dev, torch.int32).contiguous()
        xyz = points_3d.to(dev, torch.float32).contiguous()
        updated = points_ins_ids.to(dev, torch.int32).reshape(-1).clone()        # ovo.py:228
        c2w_np = c2w.detach().float().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w, np.float32)
        K_np = self.cam_intrinsics.detach().float().cpu().numpy()
        slot = kf_id % self.semmap.n_slots
        if self.sharded and self._world > 1:
            # this rank's points vote, the tables are summed over the shards, every rank decides alike (ovo_map_vote / ovo_map_apply)
            import torch.distributed as dist
            table = self.semmap.vote(xyz, updated, depth_d, seg_d, c2w_np, K_np, n_ins=self.next_ins_id, n_masks=int(binary_maps.shape[0]),
                                     match_th=self.config["match_distance_th"], track_th=int(self.config["track_th"]),
                                     depth_filter=self.config.get("depth_filter", False), rgb_depth_ratio=rgb_depth_ratio, kf_slot=slot)
            dist.all_reduce(table, op=dist.ReduceOp.SUM, group=self._group)
            votes, n_matched, self.next_ins_id = self.semmap.apply(table, self.next_ins_id)
        else:
            votes, n_matched, self.next_ins_id = self.semmap.associate(
                xyz, updated, depth_d, seg_d, c2w_np, K_np, self.next_ins_id, match_th=self.config["match_distance_th"],
                track_th=int(self.config["track_th"]), depth_filter=self.config.get("depth_filter", False),
                rgb_depth_ratio=rgb_depth_ratio, kf_slot=slot, n_masks=int(binary_maps.shape[0]))
        # (the reference loops to seg_map.max()+1 <= len(binary_maps); masks absent from seg_map get no votes, so
        # using the mask count instead saves a device->host sync without changing any result)
        n_masks = len(votes["ins_id"])

        # points that received an id in this keyframe, per mask (only materialised when someone will read it)
        new_pts = self._new_points_per_mask(slot, points_ids, points_ins_ids, votes) if self.debug_info else None

        matched_ins_info: Dict[int, List[Tuple[int, int]]] = {}          # ovo.py:254-280
        for m in range(n_masks):
            ins = int(votes["ins_id"][m])
            if ins < 0:
                continue
            area = int(votes["area"][m])
            pts = new_pts[m] if new_pts is not None else []
            if votes["is_new"][m]:
                obj = Instance3D(ins, kf_id=kf_id, points_ids=pts, mask_area=area)
                obj.bank_row = self._alloc_bank_row()
                self.objects[ins] = obj
                self._rows_cache = None
                matched_ins_info[ins] = [(m, area)]
            else:
                self.objects[ins].update(pts, kf_id, area)
                matched_ins_info.setdefault(ins, []).append((m, area))

        matched_ins_ids, maps, mask_row = self._fuse_masks_with_same_ins_id(binary_maps, matched_ins_info, kf_id, n_masks)
        if self.debug_info:
            ins_maps = torch.full(image.shape[:2], -1, dtype=torch.int32, device=dev)
            for j, ins in enumerate(matched_ins_ids):
                ins_maps[maps[j].bool()] = ins
            self.keyframes["ins_maps"].append(ins_maps.cpu().numpy())
        if self.dense:
            self._ensure_dense(xyz.shape[0])
        return matched_ins_ids, maps, n_matched, updated, dict(slot=slot, mask_row=mask_row)

    def _new_points_per_mask(self, slot, points_ids, ins_before, votes):
        """ids of the points that were unassigned before this keyframe, grouped by the mask they matched
        (what ovo.py:261 builds with a per-mask `.cpu().tolist()`): one D2H of the match list instead."""
        n = int(votes["n_matched"].sum())
        n_masks = len(votes["ins_id"])
        pairs = self.semmap.matches(slot, n)
        before = ins_before.reshape(-1).to(self._dev)
        fresh = pairs[before[pairs[:, 0].long()] == -1]                      # only points without an id count (ovo.py:261,274)
        ids = points_ids.reshape(-1).to(self._dev)[fresh[:, 0].long()].cpu().numpy()
        fresh = fresh.cpu().numpy()
        order = np.lexsort((fresh[:, 0], fresh[:, 1]))                        # by mask, then point order (as boolean indexing yields it)
        bounds = np.searchsorted(fresh[order, 1], np.arange(n_masks + 1))
        ids = ids[order]
        return [ids[bounds[m]: bounds[m + 1]].tolist() for m in range(n_masks)]

    def _fuse_masks_with_same_ins_id(self, binary_maps, matched_ins_info, kf_id, n_masks):
        """ovo.py:284-324.  Masks voted to the same instance are OR-ed (one kernel for all groups); returns
        (matched_ins_ids, maps uint8 [M',H,W], mask_row i32 [n_masks]: output row of every input mask or -1)."""
        dev = self._dev
        group = np.full(max(n_masks, binary_maps.shape[0]), -1, np.int32)
        for r, (ins, lst) in enumerate(matched_ins_info.items()):
            for m, _ in lst:
                group[m] = r
        R = len(matched_ins_info)
        if R == 0:
            return [], binary_maps[:0].to(dev, torch.uint8), torch.from_numpy(group[:n_masks]).to(dev)
        masks_u8 = binary_maps.to(dev).to(torch.uint8).contiguous()
        group_d = torch.from_numpy(group[: masks_u8.shape[0]].copy()).to(dev)
        merged, areas_d = self.semmap.merge_masks(masks_u8, group_d, R)
        multi = any(len(lst) > 1 for lst in matched_ins_info.values())
        areas = areas_d.cpu().numpy() if (multi and self.n_top_views > 0) else None
        matched_ins_ids, keep = [], []
        for r, (ins, lst) in enumerate(list(matched_ins_info.items())):
            if len(lst) > 1 and self.n_top_views > 0:
                self.objects[ins].add_top_kf(kf_id, int(areas[r]))
            if self.n_top_views <= 0 or self.objects[ins].is_top_kf(kf_id):
                matched_ins_ids.append(ins)
                keep.append(r)
            else:
                matched_ins_info.pop(ins)
        row_of_group = np.full(R, -1, np.int32)
        row_of_group[keep] = np.arange(len(keep), dtype=np.int32)
        mask_row = np.where(group[:n_masks] >= 0, row_of_group[np.clip(group[:n_masks], 0, None)], -1).astype(np.int32)
        maps = merged if len(keep) == R else merged[torch.as_tensor(keep, device=dev, dtype=torch.long)]
        return matched_ins_ids, maps, torch.from_numpy(mask_row).to(dev)

    # ------------------------------------------------------------------------------------------ keyframe: descriptors
    def compute_semantic_info(self) -> None:
        """ovo.py:326-328.  With `clip.batch_keyframes: B` (default 1 = the reference's cadence) descriptors are
        computed for B queued keyframes at a time: one ViT pass over 2B images instead of B passes over 2."""
        delay, B = self.config.get("kf_queue_delay", 0), max(1, int(self.config["clip"].get("batch_keyframes", 1)))
        if B == 1:
            if len(self.keyframes_queue) > delay:
                self._compute_semantic_batch(1)
        else:
            while len(self.keyframes_queue) - delay >= B:
                self._compute_semantic_batch(B)

    def complete_semantic_info(self) -> None:
        B = max(1, int(self.config["clip"].get("batch_keyframes", 1)))
        while len(self.keyframes_queue) > 0:
            self._compute_semantic_batch(min(B, len(self.keyframes_queue)))

    def _compute_semantic_info(self) -> None:
        self._compute_semantic_batch(1)

    def _compute_semantic_batch(self, n: int) -> None:
        """ovo.py:334-364 for the first n keyframes of the queue (encoded together, bookkeeping in queue order)."""
        items = []
        for _ in range(n):
            matched_ins_ids, binary_maps, image, kf_id, extra = self.keyframes_queue.popleft()
            if len(matched_ins_ids) == 0:
                continue
            mask_row = extra["mask_row"]
            if self.n_top_views > 0:
                sel = [j for j, ins in enumerate(matched_ins_ids) if self.objects[ins].is_top_kf(kf_id)]
                if len(sel) == 0:
                    continue
                if len(sel) != len(matched_ins_ids):
                    remap = np.full(len(matched_ins_ids), -1, np.int32)
                    remap[sel] = np.arange(len(sel), dtype=np.int32)
                    remap_d = torch.from_numpy(remap).to(self._dev)
                    mask_row = torch.where(mask_row >= 0, remap_d[mask_row.clamp_min(0).long()], mask_row)
                    matched_ins_ids = [matched_ins_ids[j] for j in sel]
                    binary_maps = binary_maps[torch.as_tensor(sel, device=self._dev, dtype=torch.long)]
            items.append((matched_ins_ids, binary_maps, image, kf_id, extra["slot"], mask_row))
        if not items:
            return
        same_shape = all(it[2].shape == items[0][2].shape for it in items)
        groups = [items] if same_shape else [[it] for it in items]
        self._enc_stream.wait_stream(torch.cuda.current_stream(self._dev))     # masks / match lists are ready
        with torch.cuda.stream(self._enc_stream):
            self._compute_groups(groups)

    def _compute_groups(self, groups) -> None:
        for group in groups:
            for it in group:                # allocated on the caller's stream, read (and released) under the descriptor stream
                it[1].record_stream(self._enc_stream)
                it[5].record_stream(self._enc_stream)
            rows_per_kf = self._extract_clip_batch([it[2] for it in group], [it[1] for it in group], [it[3] for it in group])
            if self.dense and len(group) > 1:
                # all keyframes of the batch in ONE pass over the dense bank (bit-identical to one pass per keyframe)
                base, total = rows_per_kf[0][0], sum(len(r) for r in rows_per_kf)
                nm = max(int(it[5].shape[0]) for it in group)
                mr = torch.full((len(group), nm), -1, dtype=torch.int32, device=self._dev)
                for k, (it, rows) in enumerate(zip(group, rows_per_kf)):
                    loc = it[5]
                    mr[k, : loc.shape[0]] = torch.where(loc >= 0, loc + (rows[0] - base), loc)
                self.semmap.fuse_dense_batch([it[4] for it in group], self._dense_bank, self._dense_bank_lo, self._dense_counts,
                                             self._store[base: base + total], mr)
            for (ids, _, _, kf_id, slot, mask_row), rows in zip(group, rows_per_kf):
                self._update_matched_objects_clip(rows, ids, kf_id)
                if self.dense and len(group) == 1:
                    self.semmap.fuse_dense(slot, self._dense_bank, self._dense_bank_lo, self._dense_counts,
                                           self._store[rows[0]: rows[0] + len(ids)], mask_row.contiguous())
                if self.config.get("log", False):
                    frame_id = self.keyframes["frame_id"][kf_id]
                    self.logger.log_ovo_stats({"frame_id": frame_id, "t_clip": round(self._time_cache[0], 2),
                                               "t_up": round(self._time_cache[-1], 3)}, print_output=True)
            self._time_cache = []

    def _extract_clip_batch_sharded(self, images, maps, counts, kf_ids) -> List[List[int]]:
        """Keyframe k's descriptors are computed by rank k % world (the encoder is replicated, frames are data parallel) and
        broadcast to the other ranks (<= 0.2 MB per keyframe); every rank appends them to its replica of the descriptor store."""
        import torch.distributed as dist
        D = self._store.shape[1]
        mine = [j for j, k in enumerate(kf_ids) if k % self._world == self._rank and counts[j] > 0]
        feats = [torch.empty(c, D, device=self._dev, dtype=torch.float32) for c in counts]
        if mine:
            if self.clip_generator.embed_type == "TextRegion" and all(images[j].shape == images[mine[0]].shape for j in mine):
                img = torch.stack([torch.from_numpy(np.ascontiguousarray(images[j])) for j in mine]).to(self._dev, non_blocking=True)
                out = self.clip_generator.encoder.encode_regions(img, torch.cat([maps[j] for j in mine]), masks_per_frame=[counts[j] for j in mine])
                off = 0
                for j in mine:
                    feats[j] = out[off: off + counts[j]].contiguous()
                    off += counts[j]
            else:
                for j in mine:
                    im = torch.from_numpy(np.ascontiguousarray(images[j])).to(self._dev)
                    feats[j] = self.clip_generator.extract_clip(im, maps[j]).contiguous()
        out_rows, r0 = [], self._store_n
        for j, k in enumerate(kf_ids):
            if counts[j] > 0:
                dist.broadcast(feats[j], src=dist.get_global_rank(self._group, k % self._world) if self._group is not None else k % self._world,
                               group=self._group)
                self._store[r0: r0 + counts[j]].copy_(feats[j])
            out_rows.append(list(range(r0, r0 + counts[j])))
            r0 += counts[j]
        self._store_n = r0
        return out_rows

    @profil
    def _extract_clip_batch(self, images: List[np.ndarray], maps: List[torch.Tensor], kf_ids: List[int] | None = None) -> List[List[int]]:
        """ovo.py:426-437 for a batch of keyframes: descriptors are written straight into the device descriptor
        store; returns the store rows per keyframe."""
        counts = [int(m.shape[0]) for m in maps]
        M = sum(counts)
        self._grow_store(self._store_n + M)
        if self.sharded and self._world > 1:
            return self._extract_clip_batch_sharded(images, maps, counts, kf_ids)
        if len(images) == 1:
            img = torch.from_numpy(np.ascontiguousarray(images[0])).to(self._dev, non_blocking=True)[None]
        else:
            img = torch.stack([torch.from_numpy(np.ascontiguousarray(im)) for im in images]).to(self._dev, non_blocking=True)
        masks = maps[0] if len(maps) == 1 else torch.cat(maps)
        if self.clip_generator.embed_type == "TextRegion":
            feats = self.clip_generator.encoder.encode_regions(img, masks, masks_per_frame=counts)
        else:       # crop-based types: every keyframe is already a batch of 2M+1 images (clip_generator.py:136-158)
            feats = torch.cat([self.clip_generator.extract_clip(img[f], maps[f]) for f in range(len(images))])
        self._store[self._store_n: self._store_n + M].copy_(feats)
        out, r0 = [], self._store_n
        for c in counts:
            out.append(list(range(r0, r0 + c)))
            r0 += c
        self._store_n += M
        return out

    @profil
    def _extract_clip(self, image: np.ndarray, binary_maps: torch.Tensor) -> List[int]:
        """ovo.py:426-437: one descriptor per mask; written straight into the device descriptor store.
        Returns the store rows."""
        M = binary_maps.shape[0]
        self._grow_store(self._store_n + M)
        img = torch.from_numpy(np.ascontiguousarray(image)).to(self._dev)
        feats = self.clip_generator.extract_clip(img, binary_maps)
        self._store[self._store_n: self._store_n + M].copy_(feats)
        rows = list(range(self._store_n, self._store_n + M))
        self._store_n += M
        return rows

    @profil
    def _update_matched_objects_clip(self, rows: List[int], matched_ins_ids: List[int], kf_id: int) -> None:
        """ovo.py:439-461."""
        ins_embeds = {ins: rows[i] for i, ins in enumerate(matched_ins_ids) if ins != -1}
        self.keyframes["ins_descriptors"][kf_id] = ins_embeds
        work, inc = [], []
        avg = Instance3D.mv_fusion == "avg_pooling"
        for i, ins in enumerate(matched_ins_ids):
            obj = self.objects[ins]
            obj.pending_rows.append(rows[i])
            if not obj.to_update:
                continue                      # instance3d.py:168: nothing is fused until the flag is raised again
            # avg_pooling, and the only change since the last fusion are descriptors that arrived since: the mean is updated in
            # place (O(new views)) instead of re-reading every stored view of the instance (the reference re-stacks them all,
            # instance3d.py:157-189; the same mean up to f32 rounding)
            if avg and obj.n_fused > 0 and not obj.evicted and obj._desc_epoch == self._desc_epoch:
                inc.append((obj, obj.n_fused, obj.pending_rows))
                obj.n_fused += len(obj.pending_rows)
                obj.pending_rows = []
                obj.to_update = False
                continue
            views = obj.views_to_fuse(self.keyframes["ins_descriptors"])
            if views is not None:
                obj._desc_epoch = self._desc_epoch
                work.append((obj, views))
        self._fuse(work)
        if inc:
            quads, idx = [], []
            for o, n_before, pend in inc:
                quads.append((o.bank_row, n_before, len(idx), len(idx) + len(pend)))
                idx.extend(pend)
            buf = torch.tensor([v for q in quads for v in q] + idx, dtype=torch.int32).to(self._dev, non_blocking=True)
            self.semmap.bank_add_views(self._bank, self._store, buf[: 4 * len(quads)].view(-1, 4), buf[4 * len(quads):])
            for o, _, _ in inc:
                if o.clip_feature is None or o.clip_feature.dim() == 1:     # instance3d.py:186-187: a fused feature is [1, D]
                    o.clip_feature = self._bank[o.bank_row][None]
                o.clip_feature_kf = None

    def update_objects_clip(self, force_update: bool = False) -> None:
        """ovo.py:463-470."""
        work = []
        for obj in self.objects.values():
            views = obj.views_to_fuse(self.keyframes["ins_descriptors"], force_update=force_update)
            if views is not None:
                work.append((obj, views))
        self._fuse(work)

    def _fuse(self, work) -> None:
        """One `ovo_fuse_views` launch for every (instance, views) pair (Instance3D.update_clip batched)."""
        if not work:
            return
        idx, off, out_rows = [], [0], []
        for obj, views in work:
            idx.extend(views)
            off.append(len(idx))
            out_rows.append(obj.bank_row)
        dev = self._dev
        mode = FUSION_MODES[Instance3D.mv_fusion]
        chosen = torch.zeros(len(work), device=dev, dtype=torch.int32) if mode != 0 else None
        self.semmap.fuse_views(self._store, torch.tensor(idx, dtype=torch.int32, device=dev),
                               torch.tensor(off, dtype=torch.int32, device=dev), mode, self._bank,
                               torch.tensor(out_rows, dtype=torch.int32, device=dev), chosen)
        for j, (obj, views) in enumerate(work):
            row = self._bank[obj.bank_row]
            if len(views) == 1:                                  # instance3d.py:184-185
                obj.clip_feature, obj.clip_feature_kf = row, 0
            else:                                                # instance3d.py:186-187: fused shape is [1, D]
                obj.clip_feature = row[None]
                obj.clip_feature_kf = None if chosen is None else (chosen, j)      # read back lazily (Instance3D.clip_feature_kf)

    # ------------------------------------------------------------------------------------------ tables
    def _grow_store(self, need: int) -> None:
        if need > self._store.shape[0]:
            new = torch.zeros(max(need, 2 * self._store.shape[0]), self._store.shape[1], device=self._dev)
            new[: self._store_n] = self._store[: self._store_n]
            self._store.record_stream(torch.cuda.current_stream(self._dev))   # the old table may belong to another stream's pool
            self._store = new

    def _alloc_bank_row(self) -> int:
        if self._bank_n == self._bank.shape[0]:
            self._sync_descriptors()
            new = torch.zeros(2 * self._bank.shape[0], self._bank.shape[1], device=self._dev)
            new[: self._bank_n] = self._bank
            self._bank = new
            for o in self.objects.values():                      # re-point the views
                if o.clip_feature is not None:
                    o.clip_feature = new[o.bank_row] if o.clip_feature.dim() == 1 else new[o.bank_row][None]
        self._bank_n += 1
        return self._bank_n - 1

    def _ensure_dense(self, n_points: int) -> None:
        cap = int(self.config.get("dense_capacity", 0)) or n_points
        if self._dense_bank is None:
            D = self.clip_generator.clip_dim
            self._dense_bank = torch.zeros(max(cap, n_points), D, device=self._dev, dtype=torch.bfloat16)
            self._dense_bank_lo = torch.zeros(max(cap, n_points), D, device=self._dev, dtype=torch.bfloat16)
            self._dense_counts = torch.zeros(max(cap, n_points), device=self._dev, dtype=torch.int32)
        elif n_points > self._dense_bank.shape[0]:
            # fusions of earlier keyframes may still be queued on the descriptor stream with pointers into the old bank: the copy
            # (and the release of the old tensors, which belong to this stream's pool) must come after them
            self._sync_descriptors()
            grow = max(n_points, int(1.5 * self._dense_bank.shape[0]))
            n_old = self._dense_bank.shape[0]
            nb = torch.zeros(grow, self._dense_bank.shape[1], device=self._dev, dtype=torch.bfloat16)
            nl = torch.zeros(grow, self._dense_bank.shape[1], device=self._dev, dtype=torch.bfloat16)
            nc = torch.zeros(grow, device=self._dev, dtype=torch.int32)
            nb[:n_old] = self._dense_bank
            nl[:n_old] = self._dense_bank_lo
            nc[:n_old] = self._dense_counts
            self._dense_bank, self._dense_bank_lo, self._dense_counts = nb, nl, nc

    def _sync_descriptors(self) -> None:
        """Make the current stream wait for the descriptor stream (no host sync)."""
        torch.cuda.current_stream(self._dev).wait_stream(self._enc_stream)

    def descriptors_since(self, row0: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Per-keyframe region descriptors appended to the device store since `row0` (see `_store_n`); with a pinned
        `out` the read-back is enqueued asynchronously behind the work that produces them."""
        n = self._store_n - row0
        with torch.cuda.stream(self._enc_stream):
            if out is None:
                return self._store[row0: row0 + n].clone()
            out[:n].copy_(self._store[row0: row0 + n], non_blocking=True)
            return out[:n]

    def _object_rows(self) -> torch.Tensor:
        if self._rows_cache is None or self._rows_cache.shape[0] != len(self.objects):
            self._rows_cache = torch.tensor([o.bank_row for o in self.objects.values()], dtype=torch.int32, device=self._dev)
        return self._rows_cache

    # ------------------------------------------------------------------------------------------ queries
    @torch.no_grad()
    def query(self, queries: List[str], templates: List[str] = ['{}'], ensemble: bool = False) -> torch.Tensor:
        """ovo.py:495-510: [n_obj, n_queries] similarity (objects in dict order)."""
        assert len(self.objects) > 0, "No 3D instances to query!"
        self._sync_descriptors()
        self._refresh_missing_clips()
        return self.clip_generator.get_embed_txt_similarity(self._bank, queries, templates=templates, rows=self._object_rows())

    @torch.no_grad()
    def query_points(self, queries: List[str], templates: List[str] = ['{}'], n_points: int | None = None) -> torch.Tensor:
        """Dense mode: [n_points, n_queries] similarity of every map point's running-mean feature."""
        assert self.dense and self._dense_bank is not None, "dense_map mode is off or no keyframe has been fused yet"
        self._sync_descriptors()
        if isinstance(templates, str):
            templates = [templates]
        txt = self.clip_generator.text_bank([[t.format(q) for t in templates] for q in queries])
        bank = self._dense_bank if n_points is None else self._dense_bank[:n_points]
        return self.semmap.query_dense(bank, txt)

    @torch.no_grad()
    def classify_instances(self, classes: List[str], template: str | List[str] = "This is a photo of a {}", th: float = 0):
        """ovo.py:472-492."""
        sim = self.query(classes, template)
        cls, conf = self.semmap.classify(sim, th)
        return {"classes": cls.cpu().numpy().astype(np.int64), "conf": conf.cpu().numpy()}

    def _refresh_missing_clips(self) -> None:
        work = []
        for obj in self.objects.values():                         # "This should never happen" (ovo.py:522-526)
            if obj.clip_feature is None:
                obj.to_update = True
                views = obj.views_to_fuse(self.keyframes["ins_descriptors"])
                if views is not None:
                    work.append((obj, views))
        self._fuse(work)

    @torch.no_grad()
    def get_objs_clips(self) -> torch.Tensor:
        """ovo.py:512-527: [n_obj, clip_dim] on the device."""
        self._sync_descriptors()
        self._refresh_missing_clips()
        return self._bank.index_select(0, self._object_rows().long())

    # ------------------------------------------------------------------------------------------ loop closure
    def update_map(self, map_data, kfs):
        """ovo.py:366-424 (SLAM loop closure; rare, outside the measured path): flush the queue, forget deleted
        keyframes, drop instances without points, merge instances that pass the centroid / cosine / point-distance
        test of instance_utils.same_instance (instance_utils.py:5-24), re-fuse descriptors."""
        self.complete_semantic_info()
        self._sync_descriptors()
        self._desc_epoch += 1
        points_3d, _, points_ins_ids = map_data
        for i, kf in enumerate(self.keyframes["frame_id"]):
            if kf not in kfs:
                self.keyframes["ins_descriptors"].pop(kf, None)
                self.keyframes["frame_id"][i] = "Deleted"
        present = set(points_ins_ids.unique().tolist())
        objects_list = [o for i, o in self.objects.items() if i in present]
        n_del = len(self.objects) - len(objects_list)
        pcds = {}
        for o in objects_list:
            p = points_3d[points_ins_ids.reshape(-1) == o.id]
            pcds[o.id] = (p, p.mean(dim=0))
        objects, fused = {}, {}
        for i, a in enumerate(objects_list):
            if a.id in fused:
                continue
            for b in objects_list[i + 1:]:
                if b.id in fused:
                    continue
                if self._same_instance(a, b, pcds[a.id], pcds[b.id]):
                    a.add_points_ids(b.points_ids)
                    for kf in b.kfs_ids:
                        a.add_keyframes(kf)
                    for area, kf in b.top_kf:
                        a.add_top_kf(kf, area)
                    points_ins_ids[points_ins_ids == b.id] = a.id
                    fused[b.id] = a.id
            objects[a.id] = a
        print(f"Semantic Map update: removed {n_del}, fused {len(fused)} instances")
        for id2, id1 in fused.items():
            for kf in self.objects[id2].kfs_ids:
                d = self.keyframes["ins_descriptors"].get(kf)
                if d is None or id2 not in d:
                    continue
                d[id1] = d.pop(id2)
        self.objects = objects
        self._rows_cache = None
        self.update_objects_clip()
        return points_ins_ids

    def _same_instance(self, a, b, pc_a, pc_b) -> bool:
        pa, ca = pc_a
        pb, cb = pc_b
        if ((ca - cb) ** 2).sum().sqrt() > self.th_centroid:
            return False
        # instance_utils.py:13 indexes `clip_feature[0]`: the descriptor row of a fused feature ([1, D]) but the FIRST ELEMENT of a
        # single-view feature ([D], instance3d.py:184-185), which then broadcasts.  Kept as is (do not "fix").
        cos = torch.nn.functional.cosine_similarity(a.clip_feature[0], b.clip_feature[0], dim=0)
        if cos < self.th_cossim:
            return False
        # Open3D compute_point_cloud_distance (instance_utils.py:16-22) = nearest-neighbour distance of every point of a in b
        from .eval_utils import knn
        dmin, _ = knn(pb, pa, k=1)
        p_dist = (dmin[:, 0] < self.th_points).double().mean()     # .astype(float).mean() in the reference
        return bool(p_dist > 0.5 or (cos > 0.9 and p_dist > 0.2))

    # ------------------------------------------------------------------------------------------ checkpoint
    def capture_dict(self, debug_info: bool) -> Dict[str, Any]:
        """ovo.py:529-549: same flat keys (`ins_3d_ids`, `ins3d_{id}_clip_feature`, ...), tensors on the host."""
        self._sync_descriptors()
        scene = {"ins_3d_ids": np.asarray(list(self.objects.keys()))}
        for obj in self.objects.values():
            d = obj.export(debug_info)
            k = f"ins3d_{obj.id}_clip_feature"
            if d[k] is not None:
                d[k] = d[k].detach().cpu().clone()
            scene.update(d)
        if debug_info:
            scene["frame_id"] = np.array(self.keyframes["frame_id"])
            scene["ins_map"] = np.array(self.keyframes["ins_maps"])
            for kf_id, descs in self.keyframes["ins_descriptors"].items():
                for ins_id, row in descs.items():
                    scene[f"kf_{kf_id}_ins3d_{ins_id}_clips"] = self._store[row].cpu().numpy()
        return scene

    def restore_dict(self, scene_dict: Dict[str, Any], debug_info: bool = False):
        """ovo.py:551-575."""
        for i in scene_dict["ins_3d_ids"]:
            obj = Instance3D(int(i))
            obj.restore(scene_dict, debug_info)
            obj.bank_row = self._alloc_bank_row()
            if obj.clip_feature is not None:
                f = torch.as_tensor(obj.clip_feature).to(self._dev, torch.float32)
                self._bank[obj.bank_row] = f.reshape(-1)
                obj.clip_feature = self._bank[obj.bank_row] if f.dim() == 1 else self._bank[obj.bank_row][None]
            self.objects[obj.id] = obj
        self._rows_cache = None
        self._desc_epoch += 1
        if debug_info:
            self.keyframes["frame_id"] = list(scene_dict["frame_id"])
            n_kf = len(self.keyframes["frame_id"])
            self.keyframes["ins_maps"] = [x.squeeze() for x in np.split(scene_dict["ins_map"], n_kf)]
            for i in range(n_kf):
                self.keyframes["ins_descriptors"][i] = {}
                for ins_id in self.objects.keys():
                    d = scene_dict.get(f"kf_{i}_ins3d_{ins_id}_clips", None)
                    if d is not None:
                        self._grow_store(self._store_n + 1)
                        self._store[self._store_n] = torch.as_tensor(d).to(self._dev, torch.float32).reshape(-1)
                        self.keyframes["ins_descriptors"][i][ins_id] = self._store_n
                        self._store_n += 1
