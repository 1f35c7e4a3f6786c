"""Host side of the fused tcgen05 phi+pool path (pcc_deepsets_phi_pool_fwd / _bwd).

Replaces /root/reference/models/deep_sets.py:89-106 and its autograd with two kernel
launches; per-point activations never reach HBM (the backward recomputes them per tile).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib as L
from ._lib import ACT, POOL, PhiDesc, call, ptr


def _build_desc(plan: List[dict], act: str, pooling: str, tensors=None) -> PhiDesc:
    d = PhiDesc()
    d.n_layers = len(plan)
    d.input_dim = plan[0]["lin"].in_features
    d.hidden = plan[0]["lin"].out_features
    d.act = ACT[act]
    d.pooling = POOL[pooling]
    mask = 0
    for i, Lr in enumerate(plan):
        if Lr["res"]:
            mask |= 1 << i
    d.residual_mask = mask
    if tensors is not None:
        for i, (w, b) in enumerate(tensors):
            d.w[i] = w.data_ptr()
            d.b[i] = b.data_ptr()
    return d


def phi_supported(plan: List[dict], act: str, pooling: str) -> bool:
    """Static shape/feature check (no device pointers needed)."""
    if len(plan) < 2 or len(plan) > L.MAX_PHI_LAYERS or act not in ("relu", "gelu", "silu"):
        return False
    H = plan[0]["lin"].out_features
    for i, Lr in enumerate(plan):
        lin = Lr["lin"]
        if Lr["ln"] is not None or lin.bias is None or lin.out_features != H:
            return False
        if i > 0 and lin.in_features != H:
            return False
        if i == 0 and Lr["res"]:
            return False
        if (i == len(plan) - 1) == Lr["act"]:  # hidden layers activate, the final Linear does not
            return False
    d = _build_desc(plan, act, pooling)
    lib = L.load()
    return lib.pcc_phi_fused_supported(C.byref(d)) == 0


class FusedPhiPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offsets, meta, *params):
        plan_len, act, pooling, res_mask = meta
        if ctx.needs_input_grad[0]:
            # the fused backward produces parameter gradients only (the reference never asks for dx: the collate's
            # x does not require grad); failing loudly beats returning a silent None
            raise RuntimeError("the fused bf16 phi+pool path does not produce a gradient for x; use precision='fp32' "
                               "or detach x")
        x = L.f32c(x)
        dev = L.require_cuda(x, offsets, *params)
        st = L.stream_ptr(dev)
        ws_ = [L.f32c(p) for p in params]
        tensors = [(ws_[2 * i], ws_[2 * i + 1]) for i in range(plan_len)]
        d = PhiDesc()
        d.n_layers, d.input_dim, d.hidden = plan_len, tensors[0][0].shape[1], tensors[0][0].shape[0]
        d.act, d.pooling, d.residual_mask = ACT[act], POOL[pooling], res_mask
        for i, (w, b) in enumerate(tensors):
            d.w[i], d.b[i] = w.data_ptr(), b.data_ptr()
        n, B, H = x.shape[0], offsets.numel() - 1, d.hidden
        ws_bytes = call("pcc_phi_fused_workspace_bytes", C.byref(d), n, B)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        pooled = torch.empty((B, H), dtype=torch.float32, device=x.device)
        # max: argmax rows; sum / mean: aux buffer that receives the pooled hidden activations (pooling is commuted
        # with the final Linear, include/pcc.h) and is handed back to the backward in the same slot
        arg = torch.empty((B, H), dtype=torch.int32 if pooling == "max" else torch.float32, device=x.device)
        wpack = torch.empty(call("pcc_phi_packed_bytes", C.byref(d)), dtype=torch.uint8, device=x.device)
        call("pcc_deepsets_phi_pool_fwd", C.byref(d), ptr(x), ptr(offsets), n, B, ptr(pooled), ptr(arg), ptr(ws), ptr(wpack),
             dev, st)
        ctx.save_for_backward(x, offsets, arg, wpack, *ws_)
        ctx.meta = meta
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        plan_len, act, pooling, res_mask = ctx.meta
        x, offsets, arg, wpack, *ws_ = ctx.saved_tensors
        dpooled = L.f32c(dpooled)
        dev = L.require_cuda(dpooled)
        st = L.stream_ptr(dev)
        tensors = [(ws_[2 * i], ws_[2 * i + 1]) for i in range(plan_len)]
        d = PhiDesc()
        d.n_layers, d.input_dim, d.hidden = plan_len, tensors[0][0].shape[1], tensors[0][0].shape[0]
        d.act, d.pooling, d.residual_mask = ACT[act], POOL[pooling], res_mask
        for i, (w, b) in enumerate(tensors):
            d.w[i], d.b[i] = w.data_ptr(), b.data_ptr()
        n, B, H = x.shape[0], offsets.numel() - 1, d.hidden
        if pooling == "max" and B * H < n:
            # Max pooling sends gradient only to the argmax rows (autograd of deep_sets.py:104): every
            # other row of dphi is exactly zero and contributes nothing to any dW / db.  Run the backward
            # on the B*H "virtual" rows (b, f) -> x[argmax[b, f]] instead of all n rows: set b owns the H
            # virtual rows b*H .. b*H+H-1 and feature f's argmax is virtual row b*H+f, so the same kernels
            # produce the identical sums with n/(B*H) times less work.
            from .functional import gather_rows
            # an empty set has argmax -1 and receives no gradient: its pooled-gradient entries are zeroed by the
            # same launch, so the (clamped) gathered row contributes nothing
            x, dpooled = gather_rows(x, arg, dpooled)
            # the library exploits the one-hot structure of these rows (argmax = None): no dgrad / wgrad GEMM for
            # the final Linear; offsets are not read in this mode
            arg, n = None, B * H
        from .distributed import grad_like
        grads = [grad_like(t) for t in ws_]
        dw = (C.c_void_p * L.MAX_PHI_LAYERS)(*[grads[2 * i].data_ptr() for i in range(plan_len)])
        db = (C.c_void_p * L.MAX_PHI_LAYERS)(*[grads[2 * i + 1].data_ptr() for i in range(plan_len)])
        ws_bytes = call("pcc_phi_fused_workspace_bytes", C.byref(d), n, B)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        call("pcc_deepsets_phi_pool_bwd", C.byref(d), ptr(x), ptr(offsets), n, B, ptr(dpooled), ptr(arg),
             C.cast(dw, C.c_void_p), C.cast(db, C.c_void_p), ptr(ws), ptr(wpack), dev, st)
        return (None, None, None, *grads)



def phi_pool(x, offsets, plan: List[dict], act: str, pooling: str):
    params = []
    mask = 0
    for i, Lr in enumerate(plan):
        params += [Lr["lin"].weight, Lr["lin"].bias]
        if Lr["res"]:
            mask |= 1 << i
    return FusedPhiPoolFn.apply(x, offsets, (len(plan), act, pooling, mask), *params)
