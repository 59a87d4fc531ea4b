"""What does work on a second stream cost the encoder?  The ViT step of the bench (16 images) with a synthetic kernel beside it:
HBM copy, L2-resident copy, ALU-only.  Prints ms per step for each."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict
from ovo_b200 import synth
dev = torch.device("cuda")
cfg = EncoderConfig(text_layers=0)
enc = RegionEncoder(cfg, random_state_dict(cfg, text=False), max_images=16, max_h=480, max_w=640, max_masks=512)
F = 8
rgb = torch.from_numpy(np.stack([synth.rgb(480, 640, seed=i) for i in range(F)])).to(dev)
seg, bm = synth.grid_masks(480, 640, 6, 8)
masks = torch.from_numpy(np.concatenate([bm] * F)).to(dev).to(torch.uint8)
side = torch.cuda.Stream()
big_a = torch.empty(1 << 29, dtype=torch.uint8, device=dev); big_b = torch.empty_like(big_a)      # 512 MB
small_a = torch.empty(16 << 20, dtype=torch.uint8, device=dev); small_b = torch.empty_like(small_a)  # 16 MB (L2 resident)
alu = torch.randn(1 << 20, device=dev)

def hbm():      # 4 GB of HBM traffic
    for _ in range(4):
        big_b.copy_(big_a)
def l2():       # 4 GB of L2 traffic, 16 MB working set
    for _ in range(128):
        small_b.copy_(small_a)
def alu_only():
    x = alu
    for _ in range(40):
        x = torch.sin(x) * 1.0001
    return x

def run(fn, steps=20):
    for _ in range(3):
        enc.encode_regions(rgb, masks, masks_per_frame=[48] * F)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        main = torch.cuda.current_stream()
        enc.encode_regions(rgb, masks, masks_per_frame=[48] * F)
        if fn is not None:
            with torch.cuda.stream(side):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

def alone(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

base = run(None)
print(f"encoder alone {base:.3f} ms")
for name, fn in (("hbm copy 4 GB", hbm), ("l2 copy 4 GB", l2), ("alu", alu_only)):
    print(f"{name}: alone {alone(fn):.3f} ms, beside the encoder step {run(fn):.3f} ms (+{run(fn) - base:.3f})")
