"""Fused multi-tensor Adam / AdamW: a drop-in for the optimizers the reference constructs at
/root/reference/models/wrapper.py:30-33 (`torch.optim.Adam(params, lr=...)`, `torch.optim.AdamW(params, lr=...)`),
with the whole step in ONE kernel launch (pcc_optim.cu) and the step counter on the device, so `step()` can be
captured in the CUDA graph of the train step (`GraphedTrainStep(..., optimizer=FusedAdam(...))`)."""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import call, ptr


class FusedAdam(torch.optim.Optimizer):
    """decoupled=True: torch.optim.AdamW (default weight_decay 0.01); decoupled=False: torch.optim.Adam (0.0).

    Used the way models/wrapper.py uses its optimizer: zero_grad() and step().  Not carried over from torch.optim:
    * `state_dict()` holds the hyper-parameters only — the moments live in two flat device buffers and are not
      checkpointed (the reference never saves optimizer state);
    * hyper-parameters are passed to the kernel by value, so a step captured in a CUDA graph keeps the lr it was
      captured with (the reference trains with a constant lr; re-capture after changing it);
    * one step counter per parameter group."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=None, decoupled=True):
        if weight_decay is None:
            weight_decay = 1e-2 if decoupled else 0.0
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self._groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._groups.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda" or any(p.dtype != torch.float32 or not p.is_contiguous() for p in ps):
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
            offs, off = [], 0
            for p in ps:
                offs.append(off)
                off += (p.numel() + 3) // 4 * 4
            st = dict(params=ps, offsets=offs, total=off,
                      exp_avg=torch.zeros(off, dtype=torch.float32, device=dev),
                      exp_avg_sq=torch.zeros(off, dtype=torch.float32, device=dev),
                      step=torch.zeros((), dtype=torch.int64, device=dev),
                      table=torch.zeros((len(ps), 4), dtype=torch.int64, device=dev),
                      host=torch.zeros((len(ps), 4), dtype=torch.int64).pin_memory(), key=None, copied=None, graph_hosts=[],
                      spare_hosts=[torch.zeros((len(ps), 4), dtype=torch.int64).pin_memory() for _ in range(4)])
            self._groups.append(st)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group, st in zip(self.param_groups, self._groups):
            if st is None:
                continue
            key = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0) for p in st["params"])
            if key != st["key"]:   # pointer table: rebuilt only when a gradient tensor moved (never inside a graph replay)
                rows = []
                for i, (p, off) in enumerate(zip(st["params"], st["offsets"])):
                    g = p.grad
                    if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                        raise RuntimeError("FusedAdam needs contiguous fp32 gradients")
                    rows.append((key[i][0], key[i][1], p.numel(), off))
                capturing = torch.cuda.is_current_stream_capturing()
                if st["copied"] is not None and not capturing:
                    st["copied"].synchronize()       # the previous upload has left the pinned staging rows
                host = st["host"]
                if capturing:   # a captured upload re-reads its host rows at every replay: give it rows of its own
                    # (allocated ahead of time: pinning memory is not a capturable operation)
                    if not st["spare_hosts"]:
                        raise RuntimeError("FusedAdam: more than 4 graph captures with moving gradient tensors; "
                                           "keep the gradients in a GradArena or build a new optimizer")
                    host = st["spare_hosts"].pop()
                    st["graph_hosts"].append(host)
                host.copy_(torch.tensor(rows, dtype=torch.int64))
                st["table"].copy_(host, non_blocking=True)
                if not capturing:
                    st["copied"] = torch.cuda.Event()
                    st["copied"].record()
                st["key"] = key
            dev = st["exp_avg"].device.index
            b1, b2 = group["betas"]
            call("pcc_adam_step", ptr(st["table"]), len(st["params"]), st["total"], ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]),
                 ptr(st["step"]), float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                 1 if group["decoupled"] else 0, dev, L.stream_ptr(dev))
        return loss
