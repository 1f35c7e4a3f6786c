"""Device-side collate: the batches of the reference's DataLoaders, built on the GPU.

The reference collates on the host with per-sample python loops — `_collate_sparse`
(/root/reference/utils/data.py:651-663: `cat` of the features, `idx = cat([full((n_i,), i) ...])`, stacked labels)
and `_graph_collate` (:1228-1261: additionally `edges_i + node_offset`, `membership`, optional weights) — and
`ModelWrapper` then moves every tensor with a blocking pageable `.to(device)` (wrapper.py:54-55).  At >10^5 samples/s
that is the bottleneck.  Here the per-sample arrays are concatenated once into pinned staging buffers, moved with
asynchronous copies, and everything index-like is generated on the device from the sizes:

    x, idx, labels = collate_sets(batch, device)                    # same values as _collate_sparse(batch)
    X, membership, edges, weights, y = collate_graphs(batch, device, use_weights)   # same as _graph_collate

Both are usable as `DataLoader(collate_fn=...)` replacements (the outputs already live on `device`, so the `.to`
calls of the wrapper become no-ops).  No CPU fallback: `device` must be a CUDA device.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib as L
from ._lib import call, ptr


def _as_tensor(a) -> torch.Tensor:
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))


def _pinned_cat(parts: Sequence[torch.Tensor], dim: int) -> torch.Tensor:
    """torch.cat straight into one pinned host buffer (a single asynchronous H2D copy follows)"""
    shape = list(parts[0].shape)
    shape[dim] = sum(p.shape[dim] for p in parts)
    out = torch.empty(shape, dtype=parts[0].dtype, pin_memory=True)
    torch.cat(list(parts), dim=dim, out=out)
    return out


def _offsets(sizes: Sequence[int], device) -> torch.Tensor:
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(sizes, dtype=np.int64), out=off[1:])
    return torch.from_numpy(off).pin_memory().to(device, non_blocking=True)


def _expand(offsets: torch.Tensor, n: int) -> torch.Tensor:
    dev = L.require_cuda(offsets)
    idx = torch.empty(n, dtype=torch.int64, device=offsets.device)
    call("pcc_expand_segments", ptr(offsets), offsets.numel() - 1, n, ptr(idx), dev, L.stream_ptr(dev))
    return idx


def collate_sets(batch, device):
    """[(features_i [n_i, F] float, label_i)] -> (x [sum n_i, F], idx [sum n_i] int64, labels [B, ...] float32)
    with the values of `_collate_sparse` (utils/data.py:651-663), on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pcc_b200.collate builds batches on a CUDA device (no CPU fallback)")
    feats = [_as_tensor(f) for f, _ in batch]
    labels = torch.stack([_as_tensor(l) for _, l in batch]).float()
    x = _pinned_cat(feats, 0).to(device, non_blocking=True)
    off = _offsets([f.shape[0] for f in feats], device)
    idx = _expand(off, x.shape[0])
    return x, idx, labels.pin_memory().to(device, non_blocking=True)


def collate_graphs(batch, device, use_weights: bool = False):
    """[({"features", "edges", "weights"}, label)] -> (X, membership, edges, weights | None, y [B, 1]) with the values of
    `_graph_collate` (utils/data.py:1228-1261), on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pcc_b200.collate builds batches on a CUDA device (no CPU fallback)")
    feats = [_as_tensor(g["features"]) for g, _ in batch]
    edges = [_as_tensor(g["edges"]).long() for g, _ in batch]
    y = torch.stack([_as_tensor(l) for _, l in batch]).unsqueeze(1)
    X = _pinned_cat(feats, 0).to(device, non_blocking=True)
    e_local = _pinned_cat(edges, 1).to(device, non_blocking=True)
    node_off = _offsets([f.shape[0] for f in feats], device)
    edge_off = _offsets([e.shape[1] for e in edges], device)
    membership = _expand(node_off, X.shape[0])
    dev = L.require_cuda(e_local)
    e_out = torch.empty_like(e_local)
    call("pcc_offset_edges", ptr(e_local), e_local.shape[1], ptr(edge_off), ptr(node_off), len(batch), ptr(e_out), dev,
         L.stream_ptr(dev))
    weights = None
    if use_weights:
        weights = _pinned_cat([_as_tensor(g["weights"]) for g, _ in batch], 0).to(device, non_blocking=True)
    return X, membership, e_out, weights, y.pin_memory().to(device, non_blocking=True)
