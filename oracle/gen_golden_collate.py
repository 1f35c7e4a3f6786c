"""Generate tests/golden/collate.npz with the UNMODIFIED reference collate functions
(/root/reference/utils/data.py:651-663 `_collate_sparse`, :1228-1261 `_graph_collate`).  Build container only.
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

sys.dont_write_bytecode = True
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden_edge_weights import _stub, REF  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "collate.npz")


def main():
    _stub(["h5py", "matplotlib", "seaborn", "torch_geometric"])
    sys.path.insert(0, REF)
    from utils.data import Step2PointPointCloud, Step2PointGraph  # unmodified reference classes
    rng = np.random.default_rng(11)
    out = {}
    # ---- ragged sets
    sizes = [5, 1, 130, 64, 17]
    batch = [(torch.from_numpy(rng.standard_normal((n, 6)).astype(np.float32)), torch.tensor([float(i % 2)]))
             for i, n in enumerate(sizes)]
    x, idx, labels = Step2PointPointCloud._collate_sparse(None, batch)
    for i, (f, l) in enumerate(batch):
        out[f"set_f{i}"], out[f"set_l{i}"] = f.numpy(), l.numpy()
    out["set_count"] = np.int64(len(batch))
    out["set_x"], out["set_idx"], out["set_labels"] = x.numpy(), idx.numpy(), labels.numpy()
    # ---- graphs
    gsizes = [(7, 12), (1, 0), (40, 100), (16, 31)]
    gb = []
    for i, (n, e) in enumerate(gsizes):
        g = {"features": torch.from_numpy(rng.standard_normal((n, 4)).astype(np.float32)),
             "edges": torch.from_numpy(rng.integers(0, n, size=(2, e)).astype(np.int64)),
             "weights": torch.from_numpy(rng.random(e).astype(np.float32))}
        gb.append((g, torch.tensor(float(i % 2))))
        out[f"g_f{i}"], out[f"g_e{i}"], out[f"g_w{i}"], out[f"g_l{i}"] = (g["features"].numpy(), g["edges"].numpy(),
                                                                       g["weights"].numpy(), gb[-1][1].numpy())
    out["g_count"] = np.int64(len(gb))
    for uw in (True, False):
        self_ = types.SimpleNamespace(use_weights=uw)
        X, memb, edges, w, y = Step2PointGraph._graph_collate(self_, gb)
        tag = "w" if uw else "nw"
        out[f"gc_{tag}_X"], out[f"gc_{tag}_memb"], out[f"gc_{tag}_edges"], out[f"gc_{tag}_y"] = (
            X.numpy(), memb.numpy(), edges.numpy(), y.numpy())
        if uw:
            out["gc_w_weights"] = w.numpy()
        else:
            assert w is None
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
