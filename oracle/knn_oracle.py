"""CPU oracle for the kNN graph build.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference contains no kNN at all (edges are built offline from
Geant4 ancestry, /root/reference/utils/data.py:847-929, and only offset-concatenated at
batch time, :1228-1261).  kNN is the new capability BASELINE.json's north_star names;
this numpy restatement *defines* its semantics and the CUDA kernel must reproduce it
bit-exactly:

  * per cloud (contiguous rows offsets[b]:offsets[b+1]), for every centre i the k
    nearest other points j != i by squared L2 distance over the position columns;
  * distance arithmetic is float32, evaluated as ((dx*dx + dy*dy) + dz*dz) with each
    product and sum rounded to float32 (no fused multiply-add), so that every platform
    gets the same bits;
  * ordering: ascending distance, ties broken by the lower point index (stable sort);
  * clouds with fewer than k+1 points pad the missing slots with -1;
  * edge list convention is GraphConv's (graph_net.py:73 via _graph_collate,
    data.py:1228-1261): row 0 = neighbour j (source), row 1 = centre i (target),
    centre-major order, global (batch-offset) node ids.

The Gaussian edge weight follows /root/reference/utils/data.py:835-845:
  w = exp(-d^2 / (2 sigma^2)), sigma = median(d) + 1e-6 over the edges of one graph,
  d = Euclidean length of the edge.
"""
from __future__ import annotations

import numpy as np


def knn_neighbours(pos: np.ndarray, offsets: np.ndarray, k: int):
    """pos[n,3] float32, offsets[B+1] int64 -> (nbr[n,k] int64 (-1 pad), d2[n,k] float32 (inf pad))."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n = pos.shape[0]
    nbr = np.full((n, k), -1, dtype=np.int64)
    d2o = np.full((n, k), np.inf, dtype=np.float32)
    for b in range(len(offsets) - 1):
        s, e = int(offsets[b]), int(offsets[b + 1])
        p = pos[s:e]
        m = e - s
        if m == 0:
            continue
        dx = p[:, None, 0] - p[None, :, 0]
        dy = p[:, None, 1] - p[None, :, 1]
        dz = p[:, None, 2] - p[None, :, 2]
        d2 = (dx * dx + dy * dy).astype(np.float32) + (dz * dz).astype(np.float32)
        d2 = d2.astype(np.float32)
        np.fill_diagonal(d2, np.inf)
        order = np.argsort(d2, axis=1, kind="stable")
        kk = min(k, m - 1)
        if kk > 0:
            sel = order[:, :kk]
            nbr[s:e, :kk] = sel + s
            d2o[s:e, :kk] = np.take_along_axis(d2, sel, axis=1)
    return nbr, d2o


def knn_edges(nbr: np.ndarray) -> np.ndarray:
    """nbr[n,k] -> edge_index[2,E] int64 (row0 = neighbour, row1 = centre), -1 slots dropped."""
    n, k = nbr.shape
    centre = np.repeat(np.arange(n, dtype=np.int64), k)
    flat = nbr.reshape(-1)
    keep = flat >= 0
    return np.stack([flat[keep], centre[keep]])


def gaussian_edge_weights(features: np.ndarray, edges: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """data.py:835-845 restated for ONE graph (positions = features[:,1:4])."""
    positions = features[:, 1:4]
    d = np.linalg.norm(positions[edges[0]] - positions[edges[1]], axis=1)
    sigma = np.median(d) + eps
    return np.exp(-(d ** 2) / (2 * sigma ** 2)).astype(np.float32)
