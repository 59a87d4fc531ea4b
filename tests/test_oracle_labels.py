"""Pins oracle/labels.py (label transfer: KD-tree k=5 + mode) to the reference's own match_labels_to_vtx output
(tests/golden/labels.npz, oracle/gen_golden.py:gen_labels).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import labels as OL


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "labels.npz"))


@pytest.mark.parametrize("seed", [0, 1])
def test_label_transfer_matches_reference(gold, seed):
    pts, lab, vtx = OL.synth_scene(seed=seed)
    ml, masks, ids = OL.match_labels_to_vtx(torch.from_numpy(lab), torch.from_numpy(pts), torch.from_numpy(vtx))
    assert (ml.numpy() == gold[f"mesh_labels_{seed}"]).all()
    assert (ids.numpy() == gold[f"ids_{seed}"]).all()
    assert (masks.sum(1).numpy() == gold[f"mask_sums_{seed}"]).all()


def test_unfiltered_variant_matches_reference(gold):
    pts, lab, vtx = OL.synth_scene(seed=0)
    ml, _, ids = OL.match_labels_to_vtx(torch.from_numpy(lab), torch.from_numpy(pts), torch.from_numpy(vtx), filter_unasigned=False)
    assert (ml.numpy() == gold["mesh_labels_0_unfiltered"]).all() and (ids.numpy() == gold["ids_0_unfiltered"]).all()
    assert (ml < 0).any() and (ids >= 0).all()


def test_tree_equals_exhaustive_search():
    rng = np.random.default_rng(3)
    p, q = rng.random((3000, 3)).astype(np.float32), rng.random((400, 3)).astype(np.float32)
    for k in (1, 5, 8):
        dt, it = OL.knn_tree(p, q, k)
        db, ib = OL.knn_brute(p, q, k)
        assert (it == ib).all() and np.abs(dt - db).max() < 1e-12
    assert np.allclose(OL.point_cloud_distance(q, p), OL.knn_brute(p, q, 1)[0][:, 0], atol=1e-12)
