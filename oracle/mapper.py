"""TEST INFRASTRUCTURE — CPU restatement (numpy, f32, fixed operation order) of the reference's map producer
`VanillaMapper.map` (ovo/slam/vanilla_mapper.py:46-85), SURVEY §8f rank 3.  Pinned against the reference in
tests/golden/mapper.npz.  Not part of the product."""
import numpy as np

from . import fusion as OF

f32 = np.float32


def integrate_frame(xyz: np.ndarray, depth: np.ndarray, c2w: np.ndarray, K: np.ndarray, match_th=0.03, downscale=2, k_pool=3):
    """Returns the NEW points [n,3] f32 (row-major order of the down-scaled, still-unmapped, valid-depth pixels) and
    their pixel coordinates (v, u)."""
    h, w = depth.shape
    mask = depth > 0                                               # vanilla_mapper.py:56
    if len(xyz) > 0:                                               # :58-63
        corners = OF.frustum_corners(depth, c2w, K)
        planes = OF.frustum_planes(corners)
        fm = OF.frustum_mask(xyz, corners, planes)
        w2c = np.linalg.inv(c2w.astype(np.float64)).astype(f32) if False else None
        import torch
        w2c = torch.linalg.inv(torch.from_numpy(c2w)).numpy()
        idx = np.nonzero(fm)[0]
        ok, u, v = OF.project_match(xyz[idx], depth, w2c, K, match_th)
        mask = mask.copy()
        mask[v[ok], u[ok]] = False                                 # do not project depth on points already matched
        if k_pool > 1:                                             # ~maxpool(~mask): a pixel survives iff its whole 3x3 does
            r = k_pool // 2
            pad = np.pad(mask, r, mode="constant", constant_values=True)
            out = np.ones_like(mask)
            for dy in range(k_pool):
                for dx in range(k_pool):
                    out &= pad[dy:dy + h, dx:dx + w]
            mask = out
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    ys, xs, d, m = ys[::downscale, ::downscale], xs[::downscale, ::downscale], depth[::downscale, ::downscale], mask[::downscale, ::downscale]
    ys, xs, d = ys[m], xs[m], d[m].astype(f32)
    X = ((xs.astype(f32) - f32(K[0, 2])) * d / f32(K[0, 0])).astype(f32)     # :73-75
    Y = ((ys.astype(f32) - f32(K[1, 2])) * d / f32(K[1, 1])).astype(f32)
    out = np.zeros((len(d), 3), f32)
    for r in range(3):                                             # einsum("ij,mj->mi", c2w, [X,Y,Z,1]) in a fixed order
        out[:, r] = OF._dot4(c2w[r].astype(f32), X, Y, d)
    return out, np.stack([ys, xs], 1)
