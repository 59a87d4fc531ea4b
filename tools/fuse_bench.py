"""Association batch + dense fusion of the bench workload in isolation (2M points, 8 keyframes, 48 masks, D=1024):
CUDA-event timings of ovo_map_associate_batch and ovo_map_fuse_dense_batch, HBM GB/s by algorithmic bytes.
    python tools/fuse_bench.py [points] [frames]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200 import synth  # noqa: E402
from ovo_b200.map import SemanticMap  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
sm = SemanticMap()
K = synth.intrinsics(); d0 = synth.depth_map(480, 640, 0)
xyz, ids, ins = synth.point_map(N, d0, K, synth.pose(0), seed=0)
seg, bm = synth.grid_masks(480, 640, 6, 8)
M, D = bm.shape[0], 1024
xyz_d, ins0 = torch.from_numpy(xyz).to(dev), torch.from_numpy(ins).to(dev)
seg_d = torch.from_numpy(seg).to(dev)
depth = [torch.from_numpy(synth.depth_map(480, 640, i % 4)).to(dev) for i in range(F)]
c2ws = [synth.pose(i % 4) for i in range(F)]
hi = torch.zeros(N, D, device=dev, dtype=torch.bfloat16); lo = torch.zeros_like(hi)
cnt = torch.zeros(N, device=dev, dtype=torch.int32)
feats = torch.nn.functional.normalize(torch.randn(F * M, D, device=dev), dim=-1)
mask_ins = torch.full((F, M), -1, dtype=torch.int32, device=dev)
ident = torch.arange(F * M, dtype=torch.int32, device=dev).reshape(F, M)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


state = dict(nxt=0)
ins_d = ins0.clone()


def assoc():
    v, nm, state["nxt"] = sm.associate_batch(xyz_d, ins_d, depth, [seg_d] * F, c2ws, K, state["nxt"], M, kf_slots=range(F), mask_ins_out=mask_ins)
    state["nm"] = nm


ms_a = timed(assoc)
mask_row = torch.where(mask_ins >= 0, ident, -1)
touched = int((cnt.new_zeros(1)).item())
ms_f = timed(lambda: sm.fuse_dense_batch(list(range(F)), hi, lo, cnt, feats, mask_row))
n_touched = int((cnt > 0).sum())
bytes_f = n_touched * (8.0 * D + 8)
print(f"points {N} frames {F}: associate_batch {ms_a:.3f} ms ({ms_a / F * 1e3:.1f} us/keyframe, n_matched {state['nm'][0]}), "
      f"fuse_dense_batch {ms_f:.3f} ms, touched {n_touched} points, {bytes_f / ms_f / 1e6:.0f} GB/s of algorithmic bytes")
