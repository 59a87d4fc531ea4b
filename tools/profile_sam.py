"""Short SAM-2.1 Hiera-L workload for ncu captures: one `generate` (set_image + 256-prompt decoder + AMG post-processing).
    python tools/profile_sam.py [n_generate]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovo_b200.sam import Sam2  # noqa: E402
from ovo_b200.sam_config import SamConfig, random_state_dict  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = SamConfig()
sam = Sam2(cfg, random_state_dict(cfg, seed=0), max_h=480, max_w=640, max_prompts=256)
rng = np.random.default_rng(5)
img = torch.from_numpy(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)).cuda()
prm = sam.amg_params(points_per_side=16, pred_iou_thresh=0.45, stability_score_thresh=0.4, box_nms_thresh=0.9999, nms_score_th=0.2)
for _ in range(n):
    seg, maps = sam.generate(img, prm)
torch.cuda.synchronize()
print("done", maps.shape)
