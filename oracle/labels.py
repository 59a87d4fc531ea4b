"""TEST INFRASTRUCTURE — CPU restatement of the reference's label transfer and loop-closure point distance
(SURVEY §8f rank 4).  NOT part of the product: only `tests/`, `smoke()` and bench.py's CPU legs may import it.

  eu = ovo/utils/eval_utils.py      match_labels_to_vtx :13-44 (SciPy KDTree.query(k=5) + torch.mode)
  iu = ovo/utils/instance_utils.py  same_instance :5-24 (Open3D compute_point_cloud_distance = nearest-neighbour distance)

The tree search lives in un-vendored dependencies (SciPy — pinned `scipy` of the reference environment; Open3D).
SciPy IS installed here (1.18), so `knn_tree` calls the same KDTree the reference calls; `knn_brute` restates the
published definition (k smallest Euclidean distances in float64, ascending) for small cases, and the two are
checked against each other and against the reference's own `match_labels_to_vtx` output (tests/golden/labels.npz).
"""
import numpy as np
import torch


def knn_brute(points: np.ndarray, queries: np.ndarray, k: int):
    """-> (dist f64 [Q,k], idx [Q,k]) by exhaustive search in float64; ties by point index."""
    p, q = points.astype(np.float64), queries.astype(np.float64)
    d2 = ((q[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    idx = np.lexsort((np.broadcast_to(np.arange(p.shape[0]), d2.shape), d2), axis=1)[:, :k]
    return np.sqrt(np.take_along_axis(d2, idx, axis=1)), idx


def knn_tree(points: np.ndarray, queries: np.ndarray, k: int):
    """eu:24-27 exactly as the reference calls it."""
    from scipy.spatial import KDTree
    d, i = KDTree(points).query(queries, k=k)
    return (d[:, None], i[:, None]) if k == 1 else (d, i)


def match_labels_to_vtx(points_3d_labels: torch.Tensor, points_3d: torch.Tensor, mesh_vtx: torch.Tensor, filter_unasigned=True):
    """eu:13-44."""
    if filter_unasigned:
        m = (points_3d_labels > -1).squeeze()
        points_3d_labels, points_3d = points_3d_labels[m], points_3d[m]
        assert len(points_3d_labels), "All points are unassigned"
    _, idx = knn_tree(points_3d.numpy(), mesh_vtx.numpy(), 5)
    mesh_labels = torch.mode(points_3d_labels[torch.from_numpy(idx)]).values
    ids = torch.unique(mesh_labels)
    if not filter_unasigned:
        while ids[0] < 0:
            ids = ids[1:]
    masks = mesh_labels[None].expand(len(ids), -1) == ids[:, None]
    return mesh_labels, masks, ids


def point_cloud_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """iu:16-22: distance of every point of a to its nearest neighbour in b (float64)."""
    return knn_tree(b, a, 1)[0][:, 0]


def synth_scene(n_points=60000, n_vtx=20000, n_ins=25, seed=0, frac_unassigned=0.2, frac_far=0.02):
    """A room-like cloud (points on the walls / floor of a 6 x 5 x 3 m box + clutter) with instance labels, and mesh
    vertices near it; a few vertices lie far outside the cloud (unmapped parts of the ground-truth mesh)."""
    rng = np.random.default_rng(seed)

    def surface(n):
        face = rng.integers(0, 5, n)
        u, v = rng.random(n), rng.random(n)
        p = np.zeros((n, 3))
        p[face == 0] = np.c_[u * 6, v * 5, np.zeros(n)][face == 0]
        p[face == 1] = np.c_[u * 6, np.zeros(n), v * 3][face == 1]
        p[face == 2] = np.c_[u * 6, np.full(n, 5.0), v * 3][face == 2]
        p[face == 3] = np.c_[np.zeros(n), u * 5, v * 3][face == 3]
        p[face == 4] = np.c_[np.full(n, 6.0), u * 5, v * 3][face == 4]
        return p
    pts = surface(n_points) + rng.normal(0, 0.004, (n_points, 3))
    centers = rng.random((n_ins, 3)) * [6, 5, 3]
    labels = np.argmin(((pts[:, None] - centers[None]) ** 2).sum(-1), axis=1).astype(np.int64)
    labels[rng.random(n_points) < frac_unassigned] = -1
    vtx = surface(n_vtx) + rng.normal(0, 0.002, (n_vtx, 3))
    far = rng.random(n_vtx) < frac_far
    vtx[far] += rng.normal(0, 3.0, (int(far.sum()), 3)) + [12, 0, 0]
    return pts.astype(np.float32), labels, vtx.astype(np.float32)
