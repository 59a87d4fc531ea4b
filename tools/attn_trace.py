"""Where an attention work item spends its cycles: clock64 stamps of CTA 0 (ovo_attn_trace) for one 16-image ViT layer.
    python tools/attn_trace.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200 import _lib  # noqa: E402
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict  # noqa: E402

os.environ["OVO_B200_GRAPHS"] = "0"
cfg = EncoderConfig(layers=1, text_layers=0)
enc = RegionEncoder(cfg, random_state_dict(cfg, text=False), max_images=16, max_masks=64)
px = torch.randn(16, 3, 336, 336, device="cuda")
for _ in range(3):
    enc.forward_features_from_pixels(px)
buf = torch.zeros(8 * 16 * 16, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.ovo_attn_trace(_lib.ptr(buf))
enc.forward_features_from_pixels(px)
torch.cuda.synchronize()
lib.ovo_attn_trace(None)
t = buf.cpu().numpy().reshape(8, 16, 16)
names = ["top", "S ready", "tmem_ld done", "max/rescale done", "exp+P stored", "barrier done"]
for it in range(5):
    if t[it, 15, 0] == 0:
        continue
    base = t[it, 15, 0]
    print(f"item {it}: blocks done +{t[it, 15, 1] - base}, last P.V landed +{t[it, 15, 2] - base}, item end +{t[it, 15, 3] - base} cycles")
    for j in range(10):
        if t[it, j, 0] == 0:
            continue
        d = [int(t[it, j, k + 1] - t[it, j, k]) for k in range(5)]
        mma = (int(t[it, j, 6] - t[it, j, 5]), int(t[it, j, 7] - t[it, j, 6])) if t[it, j, 6] else None
        fine = [int(t[it, j, b] - t[it, j, a]) if t[it, j, b] and t[it, j, a] else -1 for a, b in ((6, 8), (8, 10), (10, 11), (11, 12))]
        print(f"  block {j}: start +{int(t[it, j, 0] - base):6d} | wait S {d[0]:5d} | tmem_ld {d[1]:5d} | max/rescale {d[2]:5d} | exp+store {d[3]:5d} | fence+barrier {d[4]:5d}"
              f" | MMA thread: woke {mma[0]:5d} after the arrive, issue took {mma[1]:5d} = V wait {fine[0]} | 4 P.V + 2 commits {fine[1]} | K wait {fine[2]} | 4 Q.K + 2 commits {fine[3]}" if mma else "")
