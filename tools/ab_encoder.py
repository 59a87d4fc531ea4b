"""A/B timing aid: ViT-L/14 forward of 16 images (graph replay) + per-class breakdown, for the library named by
$OVO_B200_LIB (default: the in-tree build).  Run twice in the same gpurun call to compare two builds on the same box.
    OVO_B200_LIB=/path/libovo_b200_old.so python tools/ab_encoder.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200 import _lib  # noqa: E402
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict  # noqa: E402

cfg = EncoderConfig(text_layers=0)
enc = RegionEncoder(cfg, random_state_dict(cfg, text=False), max_images=16, max_masks=64)
px = torch.randn(16, 3, 336, 336, device="cuda")
for _ in range(4):
    enc.forward_features_from_pixels(px)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        enc.forward_features_from_pixels(px)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
_lib.profile_begin()
enc.forward_features_from_pixels(px)
torch.cuda.synchronize()
prof = _lib.profile_report()
print(os.environ.get("OVO_B200_LIB", "in-tree"), f"forward x16 images: {best:.3f} ms  ({16 * 349.2 / best:.0f} TFLOP/s algorithmic)",
      {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]})
