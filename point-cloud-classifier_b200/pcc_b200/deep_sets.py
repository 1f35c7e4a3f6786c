"""B200-native DeepSets — drop-in for /root/reference/models/deep_sets.py.

Same constructor kwargs (deep_sets.py:6-16), same `forward(x, idx)` (:139-146), same
`state_dict` layout: `phi` / `rho` are `nn.Sequential`s of stock parameter containers
in the order the reference builds them (:44-57, :59-72), so Sequential indices and the
`ResidualBlock.linear / .layer_norm` names match.  The containers are never *called*:
forward routes through libpcc.so (sm_100a CUDA).  No CPU fallback.

Two execution paths behind the same module:
  precision="bf16" (default, env PCC_PRECISION): phi + pooling run in ONE fused
      tcgen05/TMEM kernel (bf16 operands, fp32 accumulate; activations never reach HBM),
      backward in one fused recompute kernel.  Taken when pcc_phi_fused_supported().
  precision="fp32": fp32-grade kernels layer by layer (3xTF32 mma.sync with a hi/lo split; parity mode; also the path
      for LayerNorm-in-phi and widths the fused kernel does not take).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn

from . import functional as PF
from . import fused as FZ

_ACT_NAMES = {nn.ReLU: "relu", nn.GELU: "gelu", nn.SiLU: "silu"}


class ResidualBlock(nn.Module):
    """x + act(LN?(Linear(x)))  — parameter container (deep_sets.py:149-160)."""

    def __init__(self, dim, activation, layer_norm=False):
        super().__init__()
        self.linear = nn.Linear(dim, dim)
        self.layer_norm = nn.LayerNorm(dim) if layer_norm else nn.Identity()
        self.activation = activation

    def forward(self, x):  # pragma: no cover - containers are not called on the hot path
        raise RuntimeError("ResidualBlock is a parameter container; DeepSets.forward runs the CUDA path")


def _mlp_plan(seq: nn.Sequential) -> List[dict]:
    """Walk a phi/rho Sequential into [{lin, ln, act(bool), res(bool)}]."""
    plan, mods, i = [], list(seq), 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, ResidualBlock):
            ln = m.layer_norm if isinstance(m.layer_norm, nn.LayerNorm) else None
            plan.append({"lin": m.linear, "ln": ln, "act": True, "res": True})
            i += 1
        elif isinstance(m, nn.Linear):
            ln, act, j = None, False, i + 1
            if j < len(mods) and isinstance(mods[j], nn.LayerNorm):
                ln, j = mods[j], j + 1
            if j < len(mods) and type(mods[j]) in _ACT_NAMES:
                act, j = True, j + 1
            plan.append({"lin": m, "ln": ln, "act": act, "res": False})
            i = j
        else:
            raise RuntimeError(f"unexpected module in Sequential: {type(m).__name__}")
    return plan


class DeepSets(nn.Module):
    def __init__(self,
                 input_dim: int,
                 phi_layers: list,
                 rho_layers: list,
                 output_dim: int,
                 activation: str,
                 layer_norm: bool = True,
                 residual_block: bool = False,
                 sparse_batching: bool = True,
                 pooling: str = "sum",
                 precision: Optional[str] = None):
        super().__init__()
        # an unknown activation leaves the attribute unset in the reference (:21-26) and
        # fails with AttributeError at the first use below; same here.
        if activation == "relu":
            self.activation = nn.ReLU()
        elif activation == "gelu":
            self.activation = nn.GELU()
        elif activation == "silu":
            self.activation = nn.SiLU()

        phi, last_dim = [], input_dim
        for hidden in phi_layers:
            if residual_block and last_dim == hidden:
                phi.append(ResidualBlock(hidden, self.activation, layer_norm=layer_norm))
            else:
                phi.append(nn.Linear(last_dim, hidden))
                if layer_norm:
                    phi.append(nn.LayerNorm(hidden))
                phi.append(self.activation)
            last_dim = hidden
        phi.append(nn.Linear(last_dim, last_dim))
        self.phi_output_dim = last_dim
        self.phi = nn.Sequential(*phi)

        rho, last_dim = [], self.phi_output_dim
        for hidden in rho_layers:
            rho.append(nn.Linear(last_dim, hidden))
            if layer_norm:
                rho.append(nn.LayerNorm(hidden))
            rho.append(self.activation)
            last_dim = hidden
        rho.append(nn.Linear(last_dim, output_dim))
        self.rho = nn.Sequential(*rho)

        if pooling not in ["mean", "sum", "max"]:
            raise ValueError("pooling must be 'mean', 'sum', or 'max'")
        self.pooling = pooling
        self.sparse_batching = sparse_batching  # accepted and ignored, as in the reference (:78,141-146)

        self._act_name = activation
        self.precision = precision or os.environ.get("PCC_PRECISION", "bf16")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._phi_plan = _mlp_plan(self.phi)
        self._rho_plan = _mlp_plan(self.rho)
        self.last_path = None  # "fused-bf16" | "fp32" — which kernels the last forward used

    # ---------------------------------------------------------------- building blocks
    def _mlp(self, plan, h):
        act = self._act_name
        for L in plan:
            lin, a = L["lin"], (act if L["act"] else "none")
            if L["ln"] is not None:
                z = PF.linear_act(h, lin.weight, lin.bias, None, "none")
                h = PF.layernorm_act(z, L["ln"].weight, L["ln"].bias, h if L["res"] else None, a, L["ln"].eps)
            else:
                h = PF.linear_act(h, lin.weight, lin.bias, h if L["res"] else None, a)
        return h

    def fused_supported(self) -> bool:
        return FZ.phi_supported(self._phi_plan, self._act_name, self.pooling)

    def _forward_sparse(self, x: torch.Tensor, idx: torch.Tensor, num_sets: Optional[int] = None):
        if not x.is_cuda:
            raise RuntimeError("pcc_b200.DeepSets runs on CUDA tensors only (sm_100a kernels, no CPU fallback)")
        if num_sets is None:
            num_sets = PF.index_max(idx) + 1
        offsets = PF.segment_offsets(idx, num_sets)
        if self.precision == "bf16" and self.fused_supported():
            pooled = FZ.phi_pool(x, offsets, self._phi_plan, self._act_name, self.pooling)
            self.last_path = "fused-bf16"
        else:
            phi_x = self._mlp(self._phi_plan, x)
            pooled = PF.segment_pool(phi_x, offsets, self.pooling)
            self.last_path = "fp32"
        return self._rho(pooled)

    def _rho(self, pooled):
        """set encoder head: the tiled head kernels (one launch per layer and direction) when the stack has
        no LayerNorm and fits them (widths <= 1024); otherwise layer by layer."""
        plan = self._rho_plan
        dims = [plan[0]["lin"].in_features] + [Lr["lin"].out_features for Lr in plan]
        plain = all(Lr["ln"] is None and not Lr["res"] and Lr["lin"].bias is not None for Lr in plan)
        if plain and PF.head_supported(dims, self._act_name) and pooled.shape[0] <= 65536:
            params = []
            for Lr in plan:
                params += [Lr["lin"].weight, Lr["lin"].bias]
            return PF.mlp_head(pooled, self._act_name, params)
        return self._mlp(plan, pooled)

    def forward(self, *args, **kwargs):
        """(x[sum N_i, input_dim] f32, idx[sum N_i] i64) -> logits[B, output_dim].
        Optional keyword `num_sets=B` skips the one device->host read of max(idx)."""
        return self._forward_sparse(*args, **kwargs)
