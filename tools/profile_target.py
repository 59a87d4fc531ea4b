"""Small, short workload for `ncu --set full` captures: the dominant kernels at bench shapes.
    python tools/profile_target.py [gemm|attn|query|fuse|assoc|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200 import synth  # noqa: E402
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict, gemm_bf16  # noqa: E402
from ovo_b200.map import SemanticMap  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
os.environ["OVO_B200_GRAPHS"] = "0"
sm = SemanticMap()
if what in ("gemm", "attn", "all"):
    cfg = EncoderConfig(layers=2, text_layers=0)
    enc = RegionEncoder(cfg, random_state_dict(cfg, text=False), max_images=16, max_masks=64)
    px = torch.randn(16, 3, 336, 336, device=dev)
    for _ in range(3):
        enc.forward_features_from_pixels(px)
if what in ("query", "all"):
    bank = torch.randn(2_000_000, 1024, device=dev).bfloat16()
    text = torch.randn(20, 1024, device=dev)
    out = torch.empty(2_000_000, 20, device=dev)
    for _ in range(3):
        sm.query_dense(bank, text, out)
if what in ("fuse", "assoc", "all"):
    K = synth.intrinsics(); d = synth.depth_map(); N = 2_000_000
    xyz, ids, ins = synth.point_map(N, d, K, synth.pose(0), seed=0)
    seg, bm = synth.grid_masks()
    xyz_d, ins_d, dd, seg_d = (torch.from_numpy(a).to(dev) for a in (xyz, ins, d, seg))
    bank = torch.zeros(N, 1024, device=dev, dtype=torch.bfloat16)
    bank_lo = torch.zeros_like(bank)
    counts = torch.zeros(N, device=dev, dtype=torch.int32)
    feats = torch.randn(48, 1024, device=dev)
    mask_row = torch.arange(48, dtype=torch.int32, device=dev)
    for i in range(3):
        sm.associate(xyz_d, ins_d, dd, seg_d, synth.pose(0), K, 0 if i == 0 else 48, kf_slot=0)
        sm.fuse_dense(0, bank, bank_lo, counts, feats, mask_row)
torch.cuda.synchronize()
print("done")
