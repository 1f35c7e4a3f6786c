"""CUDA-graph captured training step for the drop-in modules.

The reference trains with an eager python loop (/root/reference/models/wrapper.py:51-74:
forward, BCEWithLogitsLoss, zero_grad, backward, optimizer.step, loss.item()).  On a B200
the fused kernels finish a step in a few hundred microseconds, so python / launch
overhead would dominate.  `GraphedTrainStep` captures forward + loss + backward
(+ gradient all-reduce + optimizer) of a fixed-shape batch once and replays it; new
batches are copied into static device buffers (from pinned host memory when given host
tensors).  Shapes are static per instance; build one instance per batch shape (sweep.py
instantiates many shapes in one process).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist

from .distributed import GradArena, PeerAllReduce, flatten_grads, unflatten_into_grads
from .functional import BCEWithLogitsLoss, bce_logits_loss_and_grad


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, example_inputs: Sequence[torch.Tensor], example_target: torch.Tensor,
                 loss_fn: Optional[Callable] = None, forward_kwargs: Optional[dict] = None,
                 optimizer: Optional[torch.optim.Optimizer] = None, allreduce: bool = False, warmup: int = 3,
                 use_graph: bool = True):
        self.model, self.optimizer = model, optimizer
        self.loss_fn = loss_fn or BCEWithLogitsLoss()  # wrapper.py:38, fused forward + gradient kernel
        self.fused_loss = type(self.loss_fn) is BCEWithLogitsLoss
        self.kw = dict(forward_kwargs or {})
        self.allreduce = allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.allreduce else 1
        self.static_in = [t.clone() for t in example_inputs]
        self.static_y = example_target.clone()
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = None
        # data parallel: gradients land in one flat arena (no flatten / unflatten copies) that is averaged in
        # place, by the library's own peer-memory all-reduce INSIDE the captured step when the ranks share a
        # node (PCC_PEER_ALLREDUCE=0 forces the NCCL call after the replay)
        self.arena = GradArena(self.params) if self.allreduce else None
        self.peer = None
        import os as _os
        if self.allreduce and self.static_y.is_cuda and _os.environ.get("PCC_PEER_ALLREDUCE", "1") != "0":
            try:
                # collective-safe: every rank runs the same collective sequence inside and either all ranks get a
                # working instance or all ranks raise (and have released what they had mapped)
                self.peer = PeerAllReduce(self.arena.numel, self.static_y.device)
            except Exception as e:  # e.g. IPC not permitted: fall back to NCCL, loudly
                import sys as _sys
                print(f"[pcc_b200] peer all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=_sys.stderr)
                self.peer = None
        self.graph = None
        self.loss = None
        self.logits = None

        # The warm-up steps run the real step (optimizer and all-reduce included) on the example batch.  They must
        # not become part of the training trajectory (the reference loop, wrapper.py:51-74, has no hidden steps), so
        # parameters, buffers (BatchNorm running statistics, num_batches_tracked) and the optimizer state are
        # snapshotted here and restored IN PLACE after the capture (the graph holds their addresses).
        snap = self._snapshot() if optimizer is not None or any(True for _ in model.buffers()) else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_body()
                self._post()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if use_graph:
            for p in self.params:
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step_body()
            torch.cuda.synchronize()
        if snap is not None:
            self._restore(snap)
            torch.cuda.synchronize()

    def _optimizer_tensors(self):
        opt = self.optimizer
        if opt is None:
            return []
        out = []
        for st in getattr(opt, "_groups", None) or []:       # FusedAdam: flat moment buffers + device step counter
            if st is not None:
                out += [st["exp_avg"], st["exp_avg_sq"], st["step"]]
        for st in opt.state.values():                         # torch.optim: per-parameter state tensors
            if isinstance(st, dict):
                out += [v for v in st.values() if torch.is_tensor(v)]
        return out

    def _snapshot(self):
        with torch.no_grad():
            model_t = list(self.model.parameters()) + list(self.model.buffers())
            return {"model": [(t, t.detach().clone()) for t in model_t],
                    "opt": {id(t): (t, t.detach().clone()) for t in self._optimizer_tensors()}}

    def _restore(self, snap):
        with torch.no_grad():
            for t, c in snap["model"]:
                t.copy_(c)
            for t in self._optimizer_tensors():
                ent = snap["opt"].get(id(t))
                if ent is not None:
                    t.copy_(ent[1])
                else:            # state created lazily by the warm-up steps: back to its initial value (zeros)
                    t.zero_()

    # one training step on the static buffers
    def _step_body(self):
        for p in self.params:
            p.grad = None
        self.logits = self.model(*self.static_in, **self.kw)
        if self.fused_loss:
            # loss and d loss / d logits come out of one launch; the backward starts at the logits, so autograd neither
            # fills a ones tensor for the scalar loss nor multiplies the saved gradient by it (2 launches less per step)
            self.loss, dlogits = bce_logits_loss_and_grad(self.logits, self.static_y)
            backward = lambda: torch.autograd.backward(self.logits, grad_tensors=dlogits)  # noqa: E731
        else:
            self.loss = self.loss_fn(self.logits, self.static_y)
            backward = self.loss.backward
        if self.arena is not None:
            with self.arena:
                backward()
            self.in_arena = self.arena.holds_all_grads()
            if not self.in_arena:   # some gradient came from a kernel outside this package: gather by copy
                self.flat = flatten_grads(self.params)
            if self.peer is not None:
                # loss is a mean over the local batch, so gradients are averaged
                bucket = self.arena.flat if self.in_arena else self._pad4(self.flat)
                self.peer.run(bucket)
                if not self.in_arena:
                    unflatten_into_grads(bucket, self.params)
                if self.optimizer is not None:
                    self.optimizer.step()
        else:
            backward()
            if self.optimizer is not None:
                self.optimizer.step()

    def _pad4(self, flat):
        if flat.numel() == self.arena.numel:
            return flat
        out = torch.zeros(self.arena.numel, dtype=flat.dtype, device=flat.device)
        out[:flat.numel()] = flat
        return out

    def _post(self):
        if self.allreduce and self.peer is None:
            # NCCL launches are kept out of the CUDA graph (capturing them hung on this stack)
            bucket = self.arena.flat if self.in_arena else self.flat
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
            bucket.mul_(1.0 / self.world)
            if not self.in_arena:
                unflatten_into_grads(bucket, self.params)
            if self.optimizer is not None:
                self.optimizer.step()

    def load(self, inputs: Sequence[torch.Tensor], target: torch.Tensor):
        """copy a batch (host pinned or device tensors) into the static buffers"""
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.static_y.copy_(target, non_blocking=True)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        self._post()
        return self.loss

    def step(self, inputs: Sequence[torch.Tensor], target: torch.Tensor) -> torch.Tensor:
        self.load(inputs, target)
        return self.run()
