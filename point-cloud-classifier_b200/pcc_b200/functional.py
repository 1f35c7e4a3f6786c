"""torch.autograd.Function wrappers around the libpcc.so entry points.

Each Function corresponds to one reference call site (cited per class); PyTorch is used
only to own device memory and to chain the backward calls.  Backward runs on the
autograd worker thread, so every call passes device + current stream explicitly.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import ACT, POOL, call, ptr


def _ctx(*tensors):
    dev = L.require_cuda(*tensors)
    return dev, L.stream_ptr(dev)


# ------------------------------------------------------------------ segment bookkeeping
def index_max(idx: torch.Tensor) -> int:
    """max(idx) — one device->host read (the reference syncs here too: counts.tolist(),
    /root/reference/models/deep_sets.py:92)."""
    idx = L.i64c(idx)
    dev, st = _ctx(idx)
    out = torch.empty(1, dtype=torch.int64, device=idx.device)
    call("pcc_index_max", ptr(idx), idx.numel(), ptr(out), dev, st)
    return int(out.item())


def segment_offsets(idx: torch.Tensor, num_sets: int) -> torch.Tensor:
    """offsets[B+1] of the contiguous split by bincount(idx) (deep_sets.py:91-92)."""
    idx = L.i64c(idx)
    dev, st = _ctx(idx)
    off = torch.empty(num_sets + 1, dtype=torch.int64, device=idx.device)
    call("pcc_segment_offsets", ptr(idx), idx.numel(), num_sets, ptr(off), dev, st)
    return off


class SegmentPoolFn(torch.autograd.Function):
    """deep_sets.py:94-106 (sum/sqrt(n) | mean | max) and PyG global_mean_pool
    (graph_net.py:92,96)."""

    @staticmethod
    def forward(ctx, x, offsets, pooling: str):
        x = L.f32c(x)
        dev, st = _ctx(x, offsets)
        n, H = x.shape
        B = offsets.numel() - 1
        pooled = torch.empty((B, H), dtype=torch.float32, device=x.device)
        arg = torch.empty((B, H), dtype=torch.int32, device=x.device) if pooling == "max" else None
        call("pcc_segment_pool_fwd", ptr(x), ptr(offsets), n, B, H, POOL[pooling], ptr(pooled), ptr(arg), dev, st)
        if arg is not None:
            ctx.save_for_backward(offsets, arg)
        else:
            ctx.save_for_backward(offsets)
        ctx.meta = (n, B, H, pooling)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        n, B, H, pooling = ctx.meta
        saved = ctx.saved_tensors
        offsets, arg = saved[0], (saved[1] if len(saved) > 1 else None)
        dpooled = L.f32c(dpooled)
        dev, st = _ctx(dpooled)
        dx = torch.empty((n, H), dtype=torch.float32, device=dpooled.device)
        call("pcc_segment_pool_bwd", ptr(dpooled), ptr(offsets), ptr(arg), n, B, H, POOL[pooling], ptr(dx), dev, st)
        return dx, None, None


def segment_pool(x, offsets, pooling: str, return_argmax: bool = False):
    if return_argmax:
        x = L.f32c(x)
        dev, st = _ctx(x, offsets)
        n, H = x.shape
        B = offsets.numel() - 1
        pooled = torch.empty((B, H), dtype=torch.float32, device=x.device)
        arg = torch.empty((B, H), dtype=torch.int32, device=x.device)
        call("pcc_segment_pool_fwd", ptr(x), ptr(offsets), n, B, H, POOL["max"], ptr(pooled), ptr(arg), dev, st)
        return pooled, arg
    return SegmentPoolFn.apply(x, offsets, pooling)


# ------------------------------------------------------------------ dense layers (fp32-grade: 3xTF32 mma.sync)
class LinearActFn(torch.autograd.Function):
    """y = residual? + act(x W^T + b + pre_add?): nn.Linear + activation (+ ResidualBlock
    add), deep_sets.py:48-53,64-68,156-160; graph_net.py lin_rel/lin_root/fc1/fc2."""

    @staticmethod
    def forward(ctx, x, w, b, pre_add, residual, act: str):
        x, w = L.f32c(x), L.f32c(w)
        b = L.f32c(b) if b is not None else None
        pre_add = L.f32c(pre_add) if pre_add is not None else None
        residual = L.f32c(residual) if residual is not None else None
        dev, st = _ctx(x, w, b, pre_add, residual)
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        need_z = act != "none"
        z = torch.empty_like(y) if need_z else None
        call("pcc_linear_fwd", ptr(x), ptr(w), ptr(b), ptr(pre_add), ptr(residual), ptr(y), ptr(z), M, N, K, ACT[act],
             0, dev, st)
        ctx.save_for_backward(x, w, z)
        ctx.meta = (M, N, K, act, b is not None, pre_add is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z = ctx.saved_tensors
        M, N, K, act, has_b, has_pre, has_res = ctx.meta
        dy = L.f32c(dy)
        dev, st = _ctx(dy)
        if act != "none":
            dz = torch.empty_like(dy)
            call("pcc_act_bwd", ptr(dy), ptr(z), ptr(dz), dy.numel(), ACT[act], dev, st)
        else:
            dz = dy
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
            call("pcc_linear_bwd_data", ptr(dz), ptr(w), None, ptr(dx), M, N, K, dev, st)
        if ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2]):
            dw = torch.empty((N, K), dtype=torch.float32, device=dy.device)
            db = torch.empty((N,), dtype=torch.float32, device=dy.device) if has_b else None
            call("pcc_linear_bwd_weight", ptr(dz), ptr(x), ptr(dw), ptr(db), M, N, K, 0, dev, st)
        dpre = dz if (has_pre and ctx.needs_input_grad[3]) else None
        dres = dy if (has_res and ctx.needs_input_grad[4]) else None
        return dx, dw, db, dpre, dres, None


def linear_act(x, w, b=None, residual=None, act: str = "none", pre_add=None):
    return LinearActFn.apply(x, w, b, pre_add, residual, act)


class LayerNormActFn(torch.autograd.Function):
    """y = residual? + act(LayerNorm(z)): deep_sets.py:50-53,65-68,153-160."""

    @staticmethod
    def forward(ctx, z, gamma, beta, residual, act: str, eps: float):
        z, gamma, beta = L.f32c(z), L.f32c(gamma), L.f32c(beta)
        residual = L.f32c(residual) if residual is not None else None
        dev, st = _ctx(z, gamma, beta, residual)
        M, H = z.shape
        y = torch.empty_like(z)
        mean = torch.empty((M,), dtype=torch.float32, device=z.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=z.device)
        call("pcc_layernorm_fwd", ptr(z), ptr(gamma), ptr(beta), ptr(residual), ptr(y), ptr(mean), ptr(rstd), M, H,
             ACT[act], float(eps), dev, st)
        ctx.save_for_backward(z, gamma, beta, mean, rstd)
        ctx.meta = (M, H, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, gamma, beta, mean, rstd = ctx.saved_tensors
        M, H, act, has_res = ctx.meta
        dy = L.f32c(dy)
        dev, st = _ctx(dy)
        dz = torch.empty_like(z)
        dgamma = torch.zeros((H,), dtype=torch.float32, device=dy.device)
        dbeta = torch.zeros((H,), dtype=torch.float32, device=dy.device)
        call("pcc_layernorm_bwd", ptr(dy), ptr(z), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(dz), ptr(dgamma),
             ptr(dbeta), M, H, ACT[act], dev, st)
        return dz, dgamma, dbeta, (dy if has_res else None), None, None


def layernorm_act(z, gamma, beta, residual=None, act: str = "none", eps: float = 1e-5):
    return LayerNormActFn.apply(z, gamma, beta, residual, act, eps)


class BatchNormTrainFn(torch.autograd.Function):
    """nn.BatchNorm1d in training mode over rows (graph_net.py:76,84,89,100); updates the
    running statistics in place like torch does."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum: float, eps: float):
        x, gamma, beta = L.f32c(x), L.f32c(gamma), L.f32c(beta)
        dev, st = _ctx(x, gamma, beta, running_mean, running_var)
        n, Cc = x.shape
        y = torch.empty_like(x)
        mean = torch.empty((Cc,), dtype=torch.float32, device=x.device)
        invstd = torch.empty((Cc,), dtype=torch.float32, device=x.device)
        call("pcc_batchnorm_fwd_train", ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(invstd),
             ptr(running_mean), ptr(running_var), n, Cc, float(momentum), float(eps), dev, st)
        ctx.save_for_backward(x, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, invstd = ctx.saved_tensors
        dy = L.f32c(dy)
        dev, st = _ctx(dy)
        n, Cc = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty((Cc,), dtype=torch.float32, device=dy.device)
        dbeta = torch.empty((Cc,), dtype=torch.float32, device=dy.device)
        call("pcc_batchnorm_bwd", ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(invstd), ptr(dx), ptr(dgamma),
             ptr(dbeta), n, Cc, dev, st)
        return dx, dgamma, dbeta, None, None, None, None


class BatchNormEvalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps: float):
        x = L.f32c(x)
        dev, st = _ctx(x, gamma, beta, running_mean, running_var)
        n, Cc = x.shape
        y = torch.empty_like(x)
        call("pcc_batchnorm_fwd_eval", ptr(x), ptr(L.f32c(gamma)), ptr(L.f32c(beta)), ptr(running_mean),
             ptr(running_var), ptr(y), n, Cc, float(eps), dev, st)
        return y

    @staticmethod
    def backward(ctx, dy):
        raise RuntimeError("BatchNorm eval-mode backward is not part of the hot path")


def batchnorm(x, bn: torch.nn.BatchNorm1d):
    if bn.training:
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        return BatchNormTrainFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                      bn.momentum if bn.momentum is not None else 0.1, bn.eps)
    return BatchNormEvalFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)


# ------------------------------------------------------------------ graph stage
class GraphCSR:
    """CSR views (by target and, lazily, by source) of one batched edge list
    (layout of /root/reference/utils/data.py:1228-1261)."""

    def __init__(self, edges: torch.Tensor, n: int):
        edges = L.i64c(edges)
        self.src = edges[0].contiguous()
        self.dst = edges[1].contiguous()
        self.n = n
        self.E = edges.shape[1]
        self._by_dst = None
        self._by_src = None

    @property
    def by_dst(self):
        if self._by_dst is None:
            self._by_dst = self._build(self.dst)
        return self._by_dst

    def _build(self, keys) -> Tuple[torch.Tensor, torch.Tensor]:
        dev, st = _ctx(keys)
        rowptr = torch.empty(self.n + 1, dtype=torch.int64, device=keys.device)
        perm = torch.empty(max(self.E, 1), dtype=torch.int32, device=keys.device)
        ws_bytes = call("pcc_csr_workspace_bytes", self.n, self.E)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=keys.device)
        call("pcc_csr_build", ptr(keys), self.E, self.n, ptr(rowptr), ptr(perm), ptr(ws), dev, st)
        return rowptr, perm

    @property
    def by_src(self):
        if self._by_src is None:
            self._by_src = self._build(self.src)
        return self._by_src


class GraphAggregateFn(torch.autograd.Function):
    """agg_i = aggr_{e: dst(e)=i} w_e * x[src(e)] — the propagate step of PyG GraphConv
    (graph_net.py:73,82)."""

    @staticmethod
    def forward(ctx, x, w, csr: GraphCSR, aggr: str):
        x = L.f32c(x)
        w = L.f32c(w) if w is not None else None
        dev, st = _ctx(x, w)
        n, Cc = x.shape
        rowptr, perm = csr.by_dst
        out = torch.empty_like(x)
        arg = torch.empty((n, Cc), dtype=torch.int32, device=x.device) if aggr == "max" else None
        call("pcc_graph_aggregate_fwd", ptr(x), ptr(csr.src), ptr(w), ptr(rowptr), ptr(perm), n, Cc, POOL[aggr],
             ptr(out), ptr(arg), dev, st)
        ctx.csr, ctx.aggr, ctx.w, ctx.arg = csr, aggr, w, arg
        return out

    @staticmethod
    def backward(ctx, g):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        g = L.f32c(g)
        dev, st = _ctx(g)
        csr = ctx.csr
        n, Cc = g.shape
        rowptr_s, perm_s = csr.by_src
        dx = torch.empty_like(g)
        call("pcc_graph_aggregate_bwd", ptr(g), ptr(csr.dst), ptr(ctx.w), ptr(rowptr_s), ptr(perm_s),
             ptr(csr.by_dst[0]), ptr(ctx.arg), n, Cc, POOL[ctx.aggr], ptr(dx), dev, st)
        return dx, None, None, None


def graph_aggregate(x, w, csr: GraphCSR, aggr: str):
    return GraphAggregateFn.apply(x, w, csr, aggr)


# ------------------------------------------------------------------ kNN
def knn(pos: torch.Tensor, offsets: torch.Tensor, k: int, with_int32: bool = False):
    """pos[n,>=3] fp32 view with unit inner stride (e.g. features[:, 1:4]) -> (nbr[n,k] i64, d2[n,k] f32)
    [, nbr32[n,k] i32 from the same launch with with_int32=True]."""
    if pos.dtype != torch.float32 or pos.stride(1) != 1:
        pos = pos.float().contiguous()
    dev, st = _ctx(pos, offsets)
    n = pos.shape[0]
    B = offsets.numel() - 1
    nbr = torch.empty((n, k), dtype=torch.int64, device=pos.device)
    d2 = torch.empty((n, k), dtype=torch.float32, device=pos.device)
    nbr32 = torch.empty((n, k), dtype=torch.int32, device=pos.device) if with_int32 else None
    call("pcc_knn", ptr(pos), pos.stride(0), ptr(offsets), n, B, int(k), ptr(nbr), ptr(d2), ptr(nbr32), dev, st)
    return (nbr, d2, nbr32) if with_int32 else (nbr, d2)


def knn_edges(nbr: torch.Tensor) -> torch.Tensor:
    dev, st = _ctx(nbr)
    n, k = nbr.shape
    ei = torch.empty((2, n * k), dtype=torch.int64, device=nbr.device)
    call("pcc_knn_edges", ptr(nbr), n, k, ptr(ei), dev, st)
    return ei


def edge_weights(pos: torch.Tensor, edges: torch.Tensor, edge_offsets: torch.Tensor, eps: float = 1e-6,
                 return_sigma: bool = False):
    """Gaussian edge weights of the reference's graph dataset (utils/data.py:835-845) for batched graphs on device:
    one sigma = median edge length + eps per graph (edges of graph g = [edge_offsets[g], edge_offsets[g+1]))."""
    pos = pos if (pos.dtype == torch.float32 and pos.stride(1) == 1) else pos.float().contiguous()
    edges = L.i64c(edges)
    edge_offsets = L.i64c(edge_offsets)
    dev, st = _ctx(pos, edges, edge_offsets)
    E, G = edges.shape[1], edge_offsets.numel() - 1
    w = torch.empty(E, dtype=torch.float32, device=pos.device)
    sigma = torch.empty(max(G, 1), dtype=torch.float32, device=pos.device)
    ws = torch.empty(max(int(call("pcc_edge_weights_workspace_bytes", E, G)), 16), dtype=torch.uint8, device=pos.device)
    call("pcc_edge_weights", ptr(pos), pos.stride(0), ptr(edges), E, ptr(edge_offsets), G, float(eps), ptr(w), ptr(sigma),
         ptr(ws), dev, st)
    return (w, sigma[:G]) if return_sigma else w


# ------------------------------------------------------------------ fused loss, row gather
def bce_logits_loss_and_grad(logits, target):
    """(mean BCE-with-logits loss, d loss / d logits) from one kernel launch; no autograd node"""
    logits, target = L.f32c(logits.detach()), L.f32c(target)
    dev, st = _ctx(logits, target)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits)
    call("pcc_bce_logits", ptr(logits), ptr(target), logits.numel(), ptr(loss), ptr(dlogits), dev, st)
    return loss, dlogits


class BCEWithLogitsFn(torch.autograd.Function):
    """nn.BCEWithLogitsLoss(reduction='mean') (wrapper.py:38) with its gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, logits, target):
        loss, dlogits = bce_logits_loss_and_grad(logits, target)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None


def bce_with_logits(logits, target):
    return BCEWithLogitsFn.apply(logits, target)


class BCEWithLogitsLoss(torch.nn.Module):
    """drop-in for nn.BCEWithLogitsLoss() (mean reduction) running one fused kernel"""

    def forward(self, logits, target):
        return bce_with_logits(logits, target)


def gather_rows(x: torch.Tensor, idx_i32: torch.Tensor, grad: Optional[torch.Tensor] = None):
    """out[i] = x[clamp(idx[i])] for int32 row ids (used by the argmax-row backward of max pooling).  With `grad`
    (one value per index) also returns grad with the entries of negative indices (empty sets) zeroed, from the
    same launch."""
    x = L.f32c(x)
    dev, st = _ctx(x, idx_i32)
    flat = idx_i32.reshape(-1)
    out = torch.empty((flat.numel(), x.shape[1]), dtype=torch.float32, device=x.device)
    gm = torch.empty_like(grad) if grad is not None else None
    call("pcc_gather_rows", ptr(x), ptr(flat), flat.numel(), x.shape[1], x.shape[0], ptr(out), ptr(grad), ptr(gm), dev, st)
    return out if grad is None else (out, gm)


# ------------------------------------------------------------------ fused rho head
def _head_desc(ws, act: str) -> "L.HeadDesc":
    d = L.HeadDesc()
    d.n_layers = len(ws) // 2
    d.dims[0] = ws[0].shape[1]
    for i in range(d.n_layers):
        d.dims[i + 1] = ws[2 * i].shape[0]
        d.w[i], d.b[i] = ws[2 * i].data_ptr(), ws[2 * i + 1].data_ptr()
    d.act = ACT[act]
    return d


def head_supported(dims, act: str) -> bool:
    """static check: 1..4 layers, widths <= 1024"""
    if len(dims) < 2 or len(dims) > 5 or act not in ("relu", "gelu", "silu", "tanh"):
        return False
    return all(1 <= v <= 1024 for v in dims)


class MLPHeadFn(torch.autograd.Function):
    """rho(pooled) of deep_sets.py:112 (Linear/act stack without LayerNorm): one tiled launch per layer and
    direction, activation / act' / bias gradient folded into the operand loads (pcc_head.cu)."""

    @staticmethod
    def forward(ctx, x, act: str, *params):
        x = L.f32c(x)
        ws = [L.f32c(p) for p in params]
        dev, st = _ctx(x, *ws)
        d = _head_desc(ws, act)
        M = x.shape[0]
        zwidth = sum(ws[2 * i].shape[0] for i in range(d.n_layers - 1))
        y = torch.empty((M, ws[-2].shape[0]), dtype=torch.float32, device=x.device)
        zsave = torch.empty((M, max(zwidth, 1)), dtype=torch.float32, device=x.device)
        call("pcc_mlp_head_fwd", C.byref(d), ptr(x), ptr(y), ptr(zsave), M, dev, st)
        ctx.save_for_backward(x, zsave, *ws)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, zsave, *ws = ctx.saved_tensors
        dy = L.f32c(dy)
        dev, st = _ctx(dy)
        d = _head_desc(ws, ctx.act)
        from .distributed import grad_like
        grads = [grad_like(t) for t in ws]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = (C.c_void_p * 4)(*[grads[2 * i].data_ptr() for i in range(d.n_layers)])
        db = (C.c_void_p * 4)(*[grads[2 * i + 1].data_ptr() for i in range(d.n_layers)])
        ws_bytes = call("pcc_mlp_head_workspace_bytes", C.byref(d), x.shape[0])
        scratch = torch.empty(max(int(ws_bytes), 4), dtype=torch.uint8, device=x.device)
        call("pcc_mlp_head_bwd", C.byref(d), ptr(x), ptr(zsave), ptr(dy), ptr(dx), C.cast(dw, C.c_void_p),
             C.cast(db, C.c_void_p), ptr(scratch), x.shape[0], dev, st)
        return (dx, None, *grads)


def mlp_head(x, act: str, params):
    return MLPHeadFn.apply(x, act, *params)


def set_dense_precision(mode: str) -> None:
    """precision of the large dense layers behind Linear / GraphConv on the fp32 path: "fp32" (default, 3xTF32 on
    the tensor cores: fp32-grade, the parity mode) or "tf32" (single TF32: ~1e-3 relative, about 3x faster GEMMs).
    Library-wide (pcc_set_dense_precision); env PCC_DENSE=tf32 selects it at import."""
    call("pcc_set_dense_precision", {"fp32": 0, "tf32": 1}[mode])
