"""CPU oracle for the DeepSets hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product (point-cloud-classifier_b200/) never does.

This is a functional restatement (no nn.Module, parameters taken from a flat
state_dict) of the reference algorithm:

  * layer plan                -> /root/reference/models/deep_sets.py:44-57 (phi), :59-72 (rho)
  * residual block            -> /root/reference/models/deep_sets.py:149-160
  * segment bookkeeping       -> /root/reference/models/deep_sets.py:91-92 (bincount + contiguous split)
  * pooling (sum/sqrt(n), mean, max first-occurrence) -> /root/reference/models/deep_sets.py:96-106
  * rho head                  -> /root/reference/models/deep_sets.py:112
  * loss                      -> /root/reference/models/wrapper.py:38 (BCEWithLogitsLoss, mean)

Parity pin: the reference has no golden vectors of its own, so this
oracle is pinned against outputs of the reference module itself, generated in the
build container by oracle/gen_golden.py (which imports /root/reference) and committed
under tests/golden/deepsets_*.npz.  tests/test_oracle_golden.py checks it.

Arithmetic is plain torch CPU tensor ops in the dtype of the inputs (float32 for the
reference-equivalent run, float64 for a high-precision truth value).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

ACTS = ("relu", "gelu", "silu")
POOLS = ("sum", "mean", "max")


def layer_plan(prefix: str, in_dim: int, hidden: List[int], out_dim: int,
               layer_norm: bool, residual_block: bool) -> List[dict]:
    """Sequential-index bookkeeping of deep_sets.py:44-57 / :59-72.

    Each entry: {"kind": "linear"|"res"|"final", "lin": key-prefix of the Linear,
                 "ln": key-prefix of the LayerNorm or None, "in": K, "out": N}
    rho never uses residual blocks (deep_sets.py:63-68), so callers pass False there.
    """
    plan, i, last = [], 0, in_dim
    for h in hidden:
        if residual_block and last == h:
            plan.append({"kind": "res", "lin": f"{prefix}.{i}.linear",
                         "ln": f"{prefix}.{i}.layer_norm" if layer_norm else None,
                         "in": last, "out": h})
            i += 1
        else:
            plan.append({"kind": "linear", "lin": f"{prefix}.{i}",
                         "ln": f"{prefix}.{i + 1}" if layer_norm else None,
                         "in": last, "out": h})
            i += 3 if layer_norm else 2
        last = h
    plan.append({"kind": "final", "lin": f"{prefix}.{i}", "ln": None, "in": last, "out": out_dim})
    return plan


def _act(name: str, z: torch.Tensor) -> torch.Tensor:
    if name == "relu":
        return torch.relu(z)
    if name == "gelu":  # nn.GELU() default: exact erf form (deep_sets.py:24)
        return 0.5 * z * (1.0 + torch.erf(z * (1.0 / math.sqrt(2.0))))
    if name == "silu":
        return z * torch.sigmoid(z)
    raise AttributeError(f"unknown activation {name!r}")  # reference: attribute never set (:21-26)


def _layer_norm(z: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = z.mean(dim=-1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)  # biased, as nn.LayerNorm
    return (z - mu) * torch.rsqrt(var + eps) * w + b


def _round_bf16_ste(t: torch.Tensor) -> torch.Tensor:
    """value rounded to bf16 (round-to-nearest-even), identity gradient (straight-through)"""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def mlp_forward(sd: Dict[str, torch.Tensor], plan: List[dict], activation: str, h: torch.Tensor,
                operand_rounding: Optional[str] = None, round_final_weight: bool = True) -> torch.Tensor:
    """operand_rounding="bf16": the STATED arithmetic of the product's bf16 precision mode — both operands of
    every Linear are rounded to bf16, products and sums stay fp32 (tensor-core bf16 x bf16 -> fp32), bias /
    activation / residual in fp32.  Same reference algorithm (deep_sets.py:44-57,149-160), different stated
    operand precision; used by the parity tests to separate "the kernel computes what it says" (tight bound)
    from "bf16 operands differ from fp32 operands" (a property of the mode: ReLU masks flip where |z| is below
    the rounding noise)."""
    q = _round_bf16_ste if operand_rounding == "bf16" else (lambda t: t)
    for L in plan:
        w = sd[L["lin"] + ".weight"]
        if L["kind"] != "final" or round_final_weight:
            w = q(w)
        z = F.linear(q(h), w, sd[L["lin"] + ".bias"])
        if L["kind"] == "final":
            h = z
            continue
        if L["ln"] is not None:
            z = _layer_norm(z, sd[L["ln"] + ".weight"], sd[L["ln"] + ".bias"])
        a = _act(activation, z)
        h = h + a if L["kind"] == "res" else a
    return h


def segment_offsets(idx: torch.Tensor) -> torch.Tensor:
    """deep_sets.py:91-92: only the histogram of idx matters; the split is contiguous."""
    counts = torch.bincount(idx)
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=idx.device)
    off[1:] = torch.cumsum(counts, 0)
    return off


def segment_pool(phi_x: torch.Tensor, offsets: torch.Tensor, pooling: str
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """deep_sets.py:96-106.  Returns (pooled[B,H], argmax[B,H] global row ids or None)."""
    B = offsets.numel() - 1
    if pooling not in POOLS:
        raise ValueError("pooling must be 'mean', 'sum', or 'max'")
    counts = (offsets[1:] - offsets[:-1]).tolist()
    chunks = torch.split(phi_x, counts, dim=0)
    pooled_list, arg_list = [], []
    start = 0
    for chunk in chunks:
        n = chunk.size(0)
        if pooling == "sum":
            pooled_list.append(chunk.sum(dim=0) / torch.sqrt(torch.tensor(n, dtype=chunk.dtype)))
        elif pooling == "mean":
            pooled_list.append(chunk.mean(dim=0))
        else:
            v, i = chunk.max(dim=0)  # first occurrence on ties
            pooled_list.append(v)
            arg_list.append(i + start)
        start += n
    pooled = torch.stack(pooled_list)
    arg = torch.stack(arg_list) if pooling == "max" else None
    return pooled, arg


def deepsets_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, idx: torch.Tensor,
                     return_aux: bool = False, phi_operand_rounding: Optional[str] = None, argmax_rows=None):
    """cfg keys = the reference ctor kwargs (deep_sets.py:6-16).  phi_operand_rounding: see mlp_forward (phi only;
    the set head stays fp32, like the product's).  argmax_rows [B,H]: evaluate max pooling at these rows."""
    ln = cfg.get("layer_norm", True)
    phi = layer_plan("phi", cfg["input_dim"], list(cfg["phi_layers"]),
                     cfg["phi_layers"][-1] if cfg["phi_layers"] else cfg["input_dim"],
                     ln, cfg.get("residual_block", False))
    H = phi[-1]["out"]
    rho = layer_plan("rho", H, list(cfg["rho_layers"]), cfg["output_dim"], ln, False)
    pooling = cfg.get("pooling", "sum")
    if pooling not in POOLS:
        raise ValueError("pooling must be 'mean', 'sum', or 'max'")
    # sum / mean in the product's bf16 mode: pooling is commuted with the final Linear, which is then applied
    # to the pooled [B,H] activations by the fp32-grade head kernel — its weight is not rounded
    phi_x = mlp_forward(sd, phi, cfg["activation"], x, phi_operand_rounding, round_final_weight=(pooling == "max"))
    offsets = segment_offsets(idx)
    if argmax_rows is not None and pooling == "max":
        arg = argmax_rows
        pooled = phi_x[arg, torch.arange(phi_x.shape[1]).expand(arg.shape[0], -1)]
    else:
        pooled, arg = segment_pool(phi_x, offsets, pooling)
    logits = mlp_forward(sd, rho, cfg["activation"], pooled)
    if return_aux:
        return logits, {"phi_x": phi_x, "pooled": pooled, "argmax": arg, "offsets": offsets}
    return logits


def deepsets_train_step(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, idx: torch.Tensor,
                        y: torch.Tensor, phi_operand_rounding: Optional[str] = None, argmax_rows=None):
    """forward + BCEWithLogitsLoss(mean) + backward (wrapper.py:58-67).

    Returns (logits, loss, grads{name: tensor}, aux).  Gradients come from torch
    autograd over the restated forward above.
    """
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, aux = deepsets_forward(leaves, cfg, x, idx, return_aux=True, phi_operand_rounding=phi_operand_rounding,
                                   argmax_rows=argmax_rows)
    loss = F.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return logits.detach(), loss.detach(), grads, {k: (v.detach() if v is not None else None) for k, v in aux.items()}


def init_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic parameters with nn.Linear-like scale (NOT the reference RNG stream;
    parity tests always pass one state_dict to both sides)."""
    g = torch.Generator().manual_seed(seed)
    ln = cfg.get("layer_norm", True)
    phi = layer_plan("phi", cfg["input_dim"], list(cfg["phi_layers"]), cfg["phi_layers"][-1], ln,
                     cfg.get("residual_block", False))
    rho = layer_plan("rho", phi[-1]["out"], list(cfg["rho_layers"]), cfg["output_dim"], ln, False)
    sd = {}
    for L in phi + rho:
        bound = 1.0 / math.sqrt(L["in"])
        sd[L["lin"] + ".weight"] = ((torch.rand(L["out"], L["in"], generator=g) * 2 - 1) * bound).to(dtype)
        sd[L["lin"] + ".bias"] = ((torch.rand(L["out"], generator=g) * 2 - 1) * bound).to(dtype)
        if L["ln"] is not None:
            sd[L["ln"] + ".weight"] = (1.0 + 0.1 * torch.randn(L["out"], generator=g)).to(dtype)
            sd[L["ln"] + ".bias"] = (0.1 * torch.randn(L["out"], generator=g)).to(dtype)
    return sd
