"""Grid-hash k-nearest-neighbour search and label transfer on the GPU, through the C ABI, against
SciPy's KD-tree (what the reference calls, eval_utils.py:24-27), the exhaustive oracle and the reference's own
match_labels_to_vtx output (tests/golden/labels.npz).  Indices are compared exactly."""
import os
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import labels as OL
from ovo_b200 import eval_utils as EU


def _clouds(seed):
    rng = np.random.default_rng(seed)
    uniform = rng.random((50000, 3)).astype(np.float32) * [8, 6, 3]
    pts, _, vtx = OL.synth_scene(seed=seed)
    clustered = (rng.normal(0, 0.05, (20000, 3)) + rng.integers(0, 4, (20000, 1)) * [1.5, 0.2, 0.1]).astype(np.float32)
    return {"uniform": (uniform.astype(np.float32), (rng.random((5000, 3)) * [8, 6, 3]).astype(np.float32)),
            "room": (pts, vtx),                                           # 2 % of the vertices lie far outside the cloud
            "clustered": (clustered, (rng.random((3000, 3)) * 6 - 1).astype(np.float32))}


@pytest.mark.parametrize("name", ["uniform", "room", "clustered"])
@pytest.mark.parametrize("k", [1, 5, 8])
def test_knn_equals_kdtree(name, k):
    p, q = _clouds(0)[name]
    d, i = EU.knn(torch.from_numpy(p), torch.from_numpy(q), k=k)
    dt, it = OL.knn_tree(p, q, k)
    assert (i.cpu().numpy() == it).all()
    assert np.abs(d.cpu().numpy() - dt).max() < 1e-12 * max(1.0, dt.max())


@pytest.mark.parametrize("cell", [0.005, 0.02, 0.5, 50.0])
def test_cell_size_does_not_change_the_result(cell):
    """Tiny cells push most queries through the exhaustive fallback, one huge cell makes the walk a single cell."""
    p, q = _clouds(1)["room"]
    q = q[:3000]
    _, i = EU.knn(torch.from_numpy(p), torch.from_numpy(q), k=5, cell_size=cell)
    assert (i.cpu().numpy() == OL.knn_tree(p, q, 5)[1]).all()


def test_small_and_degenerate_inputs():
    rng = np.random.default_rng(5)
    p = rng.random((5, 3)).astype(np.float32)
    q = rng.random((17, 3)).astype(np.float32)
    _, i = EU.knn(torch.from_numpy(p), torch.from_numpy(q), k=5)
    assert (i.cpu().numpy() == OL.knn_brute(p, q, 5)[1]).all()
    with pytest.raises(RuntimeError):
        EU.knn(torch.from_numpy(p[:3]), torch.from_numpy(q), k=5)          # fewer points than neighbours
    with pytest.raises(RuntimeError):
        EU.knn(torch.from_numpy(p), torch.from_numpy(q), k=9)
    flat = np.c_[rng.random((4000, 2)), np.zeros(4000)].astype(np.float32)  # a plane: zero extent on one axis
    qq = rng.random((500, 3)).astype(np.float32)
    _, i = EU.knn(torch.from_numpy(flat), torch.from_numpy(qq), k=5)
    assert (i.cpu().numpy() == OL.knn_tree(flat, qq, 5)[1]).all()
    dup = np.repeat(rng.random((300, 3)).astype(np.float32), 4, axis=0)     # exact duplicates: ties break on the point index
    d, i = EU.knn(torch.from_numpy(dup), torch.from_numpy(dup[::7]), k=5)
    db, ib = OL.knn_brute(dup, dup[::7], 5)
    assert (i.cpu().numpy() == ib).all() and np.abs(d.cpu().numpy() - db).max() < 1e-12
    d0, i0 = EU.knn(torch.from_numpy(p), torch.zeros(0, 3), k=1)
    assert i0.shape == (0, 1)


@pytest.mark.parametrize("seed", [0, 1])
def test_label_transfer_matches_reference(golden_dir, seed):
    g = np.load(os.path.join(golden_dir, "labels.npz"))
    pts, lab, vtx = OL.synth_scene(seed=seed)
    ml, masks, ids = EU.match_labels_to_vtx(torch.from_numpy(lab), torch.from_numpy(pts), torch.from_numpy(vtx))
    assert ml.dtype == torch.int64
    assert (ml.cpu().numpy() == g[f"mesh_labels_{seed}"]).all()                     # bit-exact labels
    assert (ids.cpu().numpy() == g[f"ids_{seed}"]).all()
    assert (masks.sum(1).cpu().numpy() == g[f"mask_sums_{seed}"]).all()
    if seed == 0:
        ml, _, ids = EU.match_labels_to_vtx(torch.from_numpy(lab).cuda(), torch.from_numpy(pts).cuda(), torch.from_numpy(vtx).cuda(),
                                            filter_unasigned=False)
        assert (ml.cpu().numpy() == g["mesh_labels_0_unfiltered"]).all() and (ids.cpu().numpy() == g["ids_0_unfiltered"]).all()
    with pytest.raises(AssertionError):
        EU.match_labels_to_vtx(torch.full((100,), -1), torch.rand(100, 3), torch.rand(10, 3))


def test_full_size_map_2m_points():
    """BASELINE-sized map: 2M points, 1M mesh vertices; a sample against the KD-tree, self-query property on the rest."""
    rng = np.random.default_rng(7)
    pts, _, _ = OL.synth_scene(n_points=2_000_000, n_vtx=10, n_ins=4, seed=2)
    _, _, vtx = OL.synth_scene(n_points=10, n_vtx=1_000_000, n_ins=4, seed=3, frac_far=0.001)
    P, Q = torch.from_numpy(pts).cuda(), torch.from_numpy(vtx).cuda()
    EU.knn(P, Q[:1000], k=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d, i = EU.knn(P, Q, k=5)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"knn 2M points x 1M queries k=5: {dt * 1e3:.1f} ms")
    sel = rng.choice(vtx.shape[0], 3000, replace=False)
    it = OL.knn_tree(pts, vtx[sel], 5)[1]
    assert (i[torch.from_numpy(sel).cuda()].cpu().numpy() == it).all()
    assert (d[:, 1:] >= d[:, :-1]).all()
    d1, i1 = EU.knn(P, P[::20], k=1)                                                # every point is its own nearest neighbour
    assert (d1 == 0).all()
    same = i1[:, 0].long() == torch.arange(0, pts.shape[0], 20, device="cuda")
    assert same.float().mean().item() > 0.999                                        # (exact duplicates resolve to the lower index)


def test_same_instance_point_distance(tmp_path):
    """OVO._same_instance (loop closure) uses the same search with k = 1 (Open3D compute_point_cloud_distance)."""
    rng = np.random.default_rng(11)
    a = rng.random((4000, 3)).astype(np.float32)
    b = (a[:3000] + rng.normal(0, 0.03, (3000, 3))).astype(np.float32)
    d, _ = EU.knn(torch.from_numpy(b), torch.from_numpy(a), k=1)
    ref = OL.point_cloud_distance(a, b)
    assert np.abs(d[:, 0].cpu().numpy() - ref).max() < 1e-12
    assert int((d[:, 0] < 0.05).sum()) == int((ref < 0.05).sum())
