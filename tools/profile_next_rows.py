"""Short workload for ncu captures of the SURVEY 8f rows: crop-based descriptors (PE-L14 geometry, 2 layers) and the
grid-hash nearest-neighbour search (2M points, 200k queries).   python tools/profile_next_rows.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OVO_B200_GRAPHS"] = "0"
import bench  # noqa: E402
from ovo_b200 import eval_utils as EU, synth  # noqa: E402
from ovo_b200.encoder import EncoderConfig, RegionEncoder, random_state_dict  # noqa: E402

cfg = EncoderConfig(layers=2, text_layers=0)
sd = random_state_dict(cfg, text=False)
enc = RegionEncoder(cfg, sd, max_images=16, max_masks=64)
enc.install_pool_head(sd, pool_heads=8)
img = torch.from_numpy(synth.rgb(seed=1)).cuda()
_, bm = synth.grid_masks(rows=2, cols=4)
masks = torch.from_numpy(bm).cuda()
for _ in range(2):
    enc.encode_crops(img, masks, "hovsg", mask_res=384)
K, xyz, ids, ins, seg, bm2 = bench.scene(2_000_000, seed=0)
rng = np.random.default_rng(3)
sel = rng.integers(0, xyz.shape[0], 200_000)
vtx = torch.from_numpy((xyz[sel] + rng.normal(0, 0.01, (200_000, 3))).astype(np.float32)).cuda()
P = torch.from_numpy(xyz).cuda()
labels = torch.from_numpy(rng.integers(0, 300, xyz.shape[0]).astype(np.int64)).cuda()
for _ in range(2):
    EU.match_labels_to_vtx(labels, P, vtx)
torch.cuda.synchronize()
print("done", EU.knn_stats())
