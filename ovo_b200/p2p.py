"""Device-side exchange of the vote tables of a sharded map (include/ovo_b200.h ovo_xchg_*): the fused alternative to one
NCCL all-reduce per keyframe.  One process per GPU of one box; every rank allocates an inbox, the 64-byte CUDA IPC handles
are all-gathered through torch.distributed (the only use of the process group here) and each rank maps its peers' inboxes,
so that the exchange kernel stores straight into peer memory over NVLink / NVSwitch."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import Frame, VoteRow, check, ptr, stream_ptr
from .map import VOTE_FIELDS


class VoteExchange:
    def __init__(self, semmap, rank: int, world: int, device, table_ints: int, slots: int = 64, group=None):
        self.semmap, self.rank, self.world, self.device = semmap, rank, world, torch.device(device)
        self.lib = _lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.ovo_xchg_create(rank, world, slots, int(table_ints), C.byref(h)), "ovo_xchg_create")
        self.handle = h
        mine = (C.c_uint8 * 64)()
        check(self.lib.ovo_xchg_ipc_handle(self.handle, mine), "ovo_xchg_ipc_handle")
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t, group=group)
        buf = (C.c_uint8 * (64 * world))(*[int(v) for h_ in allh for v in h_.cpu().tolist()])
        with torch.cuda.device(self.device):
            check(self.lib.ovo_xchg_open_peers(self.handle, buf), "ovo_xchg_open_peers")
        dist.barrier(group=group)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ovo_xchg_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __call__(self, table: torch.Tensor, f: int) -> None:
        """In-place SUM of `table` (i32, 16-byte aligned) over the ranks; f = the keyframe's slot.  When a batch is pending on
        the map handle, only the compact front of the table travels (the instance count is read on the device)."""
        nm, nins = C.c_int(0), C.c_void_p()
        if self.lib.ovo_map_batch_info(self.semmap.handle, f, C.byref(nm), C.byref(nins)) != 0:
            nm, nins = C.c_int(0), C.c_void_p()
        check(self.lib.ovo_xchg_exchange(self.handle, ptr(table), table.numel(), nins, nm.value, f, stream_ptr(self.device)),
              "ovo_xchg_exchange")

    def associate_batch(self, xyz, ins_ids, depths, seg_maps, c2ws, K, next_ins_id, n_masks, match_th=0.05, track_th=100,
                        depth_filter=True, rgb_depth_ratio=(), kf_slots=None, w2cs=None, mask_ins_out=None, depth_ranges=None):
        """SemanticMap.associate_batch on this rank's shard, every keyframe's table summed over the ranks on the device: the
        whole batch is ONE library call (no Python, no NCCL launch per keyframe)."""
        sm = self.semmap
        F = len(depths)
        frames = sm._frames(depths, seg_maps, c2ws, K, n_masks, match_th, track_th, depth_filter, rgb_depth_ratio, w2cs, depth_ranges)
        nms = [int(frames[i].n_masks) for i in range(F)]
        stride = max(max(nms), 1) if mask_ins_out is None else int(mask_ins_out.shape[1])
        rows = (VoteRow * (F * stride))()
        nxt, nm = C.c_int(next_ins_id), (C.c_int * F)()
        slots = None if kf_slots is None else (C.c_int * F)(*[int(x) for x in kf_slots])
        check(self.lib.ovo_map_associate_batch_sharded(sm.handle, self.handle, ptr(xyz), ptr(ins_ids), xyz.shape[0], frames, F, slots,
                                                       C.byref(nxt), rows, stride, nm, ptr(mask_ins_out), stream_ptr(self.device)),
              "ovo_map_associate_batch_sharded")
        arr = np.frombuffer(rows, dtype=np.int32).reshape(F, stride, 8)
        votes = [{k: arr[f, :nms[f], i].copy() for i, k in enumerate(VOTE_FIELDS)} for f in range(F)]
        return votes, [int(x) for x in nm], nxt.value
