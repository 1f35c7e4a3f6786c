"""TEST INFRASTRUCTURE ONLY — CPU oracle of the Gaussian edge weights of the reference's graph dataset.

Restates /root/reference/utils/data.py:835-845 (`Step2PointGraph._compute_weights`), applied per graph as the
dataset does (one call per event, :814):
    positions = features[:, 1:4]
    dists = ||positions[edges[0]] - positions[edges[1]]||       (np.linalg.norm, float32)
    sigma = np.median(dists) + eps                                (eps = 1e-6; float32 arithmetic)
    weights = exp(-dists**2 / (2 * sigma**2))                     (float32)
PINNED: oracle/gen_golden_edge_weights.py imports the unmodified reference module (h5py / matplotlib / seaborn /
torch_geometric stubbed: they are not installed here and are not touched by this function), calls the real
`_compute_weights` on seeded graphs and commits tests/golden/edge_weights.npz; tests/test_oracle_golden.py checks
this restatement against it bit for bit."""
import numpy as np


def compute_weights(features: np.ndarray, edges: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """one graph: features [n, >=4] float32 (xyz in columns 1:4), edges [2, E] int -> weights [E] float32"""
    positions = features[:, 1:4]
    src_pos = positions[edges[0]]
    tgt_pos = positions[edges[1]]
    dists = np.linalg.norm(src_pos - tgt_pos, axis=1)
    sigma = np.median(dists) + eps
    weights = np.exp(-(dists ** 2) / (2 * sigma ** 2))
    return np.array(weights, dtype=np.float32)


def compute_weights_batched(features: np.ndarray, edges: np.ndarray, edge_offsets: np.ndarray, eps: float = 1e-6):
    """batched layout of utils/data.py:1228-1261 (node-offset edges, graphs back to back): one sigma per graph.
    Returns (weights [E] float32, sigma [G] float32, dists [E] float32)."""
    E = edges.shape[1]
    w = np.zeros(E, dtype=np.float32)
    G = len(edge_offsets) - 1
    sig = np.full(G, np.nan, dtype=np.float32)
    pos = features[:, 1:4]
    d_all = np.linalg.norm(pos[edges[0]] - pos[edges[1]], axis=1).astype(np.float32)
    for g in range(G):
        lo, hi = int(edge_offsets[g]), int(edge_offsets[g + 1])
        if hi <= lo:
            continue
        w[lo:hi] = compute_weights(features, edges[:, lo:hi], eps)
        sig[g] = np.float32(np.median(d_all[lo:hi]) + eps)
    return w, sig, d_all
