"""Generate tests/golden/edge_weights.npz with the UNMODIFIED reference function
`Step2PointGraph._compute_weights` (/root/reference/utils/data.py:835-845).  Build container only.
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import types

sys.dont_write_bytecode = True
import numpy as np

REF = os.environ.get("PCC_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "edge_weights.npz")


def _stub(names):
    for n in names:
        try:
            __import__(n)
        except Exception:
            m = types.ModuleType(n)
            sys.modules[n] = m
            if n == "matplotlib":
                sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
            if n == "torch_geometric":
                nn = types.ModuleType("torch_geometric.nn")
                for k in ("GraphConv", "GATConv", "SAGPooling", "global_mean_pool", "global_add_pool", "global_max_pool"):
                    setattr(nn, k, object)
                sys.modules["torch_geometric.nn"] = nn


def main():
    _stub(["h5py", "matplotlib", "seaborn", "torch_geometric"])
    sys.path.insert(0, REF)
    from utils.data import Step2PointGraph  # the reference class, unmodified
    rng = np.random.default_rng(7)
    out = {}
    sizes = [(40, 8), (33, 5), (128, 20), (2, 1), (64, 20)]   # (nodes, neighbours per node): odd and even edge counts
    for gi, (n, k) in enumerate(sizes):
        feats = rng.standard_normal((n, 4)).astype(np.float32)
        feats[:, 0] = rng.random(n).astype(np.float32)
        src = rng.integers(0, n, size=n * k)
        dst = np.repeat(np.arange(n), k)
        if gi == 1:
            src, dst = src[:-1], dst[:-1]                    # odd number of edges: the median is a single element
        edges = np.stack([src, dst]).astype(np.int64)
        w = Step2PointGraph._compute_weights(feats, edges)
        out[f"features_{gi}"], out[f"edges_{gi}"], out[f"weights_{gi}"] = feats, edges, w
    out["count"] = np.int64(len(sizes))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
