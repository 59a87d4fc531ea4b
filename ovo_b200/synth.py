"""Seeded synthetic RGB-D scenes shaped like the reference's inputs: ScanNet intrinsics
(data/working/configs/ScanNet/scannet.yaml:2-10), a smooth non-degenerate depth surface, a point map with
a fraction of points on the visible surface, and a grid of masks written the way
MaskGenerator._load_masks expects (mask_generator.py:162-190).  numpy only."""
import numpy as np

SCANNET_K = np.array([[577.590698, 0.0, 318.905426], [0.0, 578.729797, 242.683609], [0.0, 0.0, 1.0]], np.float32)


def intrinsics(h=480, w=640):
    K = SCANNET_K.copy()
    K[0] *= w / 640.0
    K[1] *= h / 480.0
    return K


def pose(frame_id: int, yaw: float = 0.0) -> np.ndarray:
    """Camera-to-world: translation 0.01*frame_id along x, optional yaw (rad) about the camera's y axis."""
    c2w = np.eye(4, dtype=np.float32)
    c2w[0, 3] = 0.01 * frame_id
    if yaw != 0.0:
        c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
        c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = c, s, -s, c
    return c2w


def depth_map(h=480, w=640, frame_id=0) -> np.ndarray:
    u = np.arange(w, dtype=np.float32)[None, :] + 3.0 * frame_id
    v = np.arange(h, dtype=np.float32)[:, None]
    d = (2.0 + 0.5 * np.sin(u / 100.0) + 0.1 * np.cos(v / 70.0)).astype(np.float32)
    d[0, 0] = 1.0
    d[h - 1, w - 1] = 3.2
    d[h // 3, w // 5] = 0.0          # an invalid-depth pixel
    return d


def rgb(h=480, w=640, seed=0) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def grid_masks(h=480, w=640, rows=6, cols=8, gap=2):
    """seg_map [h,w] int32 (-1 = none) and binary_maps [M,h,w] bool: rows x cols cells with a gap."""
    seg = np.full((h, w), -1, np.int32)
    ch, cw = h // rows, w // cols
    m = 0
    for r in range(rows):
        for c in range(cols):
            seg[r * ch + gap:(r + 1) * ch - gap, c * cw + gap:(c + 1) * cw - gap] = m
            m += 1
    bmaps = seg[None] == np.arange(m, dtype=np.int32)[:, None, None]
    return seg, bmaps


def point_map(n: int, depth: np.ndarray, K: np.ndarray, c2w: np.ndarray, seed=0, frac_visible=0.25,
              noise=0.005):
    """xyz [n,3] f32, ids [n] i32, ins_ids [n] i32 (-1).  frac_visible of the points lie on the depth
    surface (+N(0,noise)); the rest are uniform in a 16 x 12 x 6 m box around the camera."""
    rng = np.random.default_rng(seed)
    h, w = depth.shape
    nv = int(n * frac_visible)
    u = rng.uniform(0, w - 1, nv).astype(np.float32)
    v = rng.uniform(0, h - 1, nv).astype(np.float32)
    d = depth[np.rint(v).astype(int), np.rint(u).astype(int)]
    z = d + rng.normal(0, noise, nv).astype(np.float32)
    x = (u - K[0, 2]) * z / K[0, 0]
    y = (v - K[1, 2]) * z / K[1, 1]
    cam = np.stack([x, y, z, np.ones_like(z)], 1)
    vis = (cam @ c2w.T)[:, :3]
    box = rng.uniform([-8, -6, -1], [8, 6, 5], (n - nv, 3))
    xyz = np.concatenate([vis, box], 0).astype(np.float32)
    xyz = xyz[rng.permutation(n)]
    return np.ascontiguousarray(xyz), np.arange(n, dtype=np.int32), np.full(n, -1, np.int32)
