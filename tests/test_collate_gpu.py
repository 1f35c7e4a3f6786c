"""GPU: device-side collate (pcc_b200/collate.py) against tests/golden/collate.npz, which holds the outputs of the
reference's own `_collate_sparse` (utils/data.py:651-663) and `_graph_collate` (:1228-1261) on the same samples
(oracle/gen_golden_collate.py).  Index / integer outputs and copied values must be identical."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))

from pcc_b200 import collate as CL  # noqa: E402

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate.npz"))


def test_collate_sets_matches_reference():
    batch = [(torch.from_numpy(Z[f"set_f{i}"]), torch.from_numpy(Z[f"set_l{i}"])) for i in range(int(Z["set_count"]))]
    x, idx, labels = CL.collate_sets(batch, "cuda:0")
    assert x.is_cuda and idx.dtype == torch.int64 and labels.dtype == torch.float32
    assert np.array_equal(x.cpu().numpy(), Z["set_x"])
    assert np.array_equal(idx.cpu().numpy(), Z["set_idx"])
    assert np.array_equal(labels.cpu().numpy(), Z["set_labels"])


@pytest.mark.parametrize("use_weights", [True, False])
def test_collate_graphs_matches_reference(use_weights):
    batch = []
    for i in range(int(Z["g_count"])):
        g = {"features": torch.from_numpy(Z[f"g_f{i}"]), "edges": torch.from_numpy(Z[f"g_e{i}"]),
             "weights": torch.from_numpy(Z[f"g_w{i}"])}
        batch.append((g, torch.from_numpy(Z[f"g_l{i}"])))
    X, memb, edges, w, y = CL.collate_graphs(batch, "cuda:0", use_weights=use_weights)
    tag = "w" if use_weights else "nw"
    assert np.array_equal(X.cpu().numpy(), Z[f"gc_{tag}_X"])
    assert memb.dtype == torch.int64 and np.array_equal(memb.cpu().numpy(), Z[f"gc_{tag}_memb"])
    assert np.array_equal(edges.cpu().numpy(), Z[f"gc_{tag}_edges"])     # incl. the graph with no edges
    assert np.array_equal(y.cpu().numpy(), Z[f"gc_{tag}_y"])
    if use_weights:
        assert np.array_equal(w.cpu().numpy(), Z["gc_w_weights"])
    else:
        assert w is None


def test_collate_refuses_cpu():
    with pytest.raises(RuntimeError):
        CL.collate_sets([(torch.zeros(2, 3), torch.zeros(1))], "cpu")
