"""Timing sweep of ovo_knn over the grid cell size on the bench scene (2M points, 1M queries)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ovo_b200 import eval_utils as EU, _lib

K, xyz, ids, ins, seg, bm = bench.scene(2_000_000, seed=0)
rng = np.random.default_rng(3)
nq = 1_000_000
sel = rng.integers(0, xyz.shape[0], nq)
vtx = torch.from_numpy((xyz[sel] + rng.normal(0, 0.01, (nq, 3))).astype(np.float32)).cuda()
P = torch.from_numpy(xyz).cuda()
for cell in [0.0] + [float(c) for c in sys.argv[1:]]:
    for rep in range(2):
        torch.cuda.synchronize()
        _lib.profile_begin()
        t0 = time.perf_counter()
        EU.knn(P, vtx, k=5, cell_size=cell, return_distance=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        prof = _lib.profile_report()
    print(f"cell {cell}: {dt*1e3:.1f} ms wall, kernels {prof['other']['ms']:.1f} ms in {prof['other']['launches']} scopes, {EU.knn_stats()}", flush=True)
