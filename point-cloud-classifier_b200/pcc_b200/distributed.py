"""Data parallelism for the drop-in modules: one process per GPU, gradients averaged with a
single flat all-reduce (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).

The reference is single-process.  Sets / graphs are independent through the
whole hot path, so the batch shards naturally: rank r trains on its own contiguous slice of
sets (`shard_sets`), and because the loss is a mean over the local batch the summed gradients
are divided by the world size.  The modules are NOT wrapped (no `module.`
prefix in state_dict keys, wrapper.py:179-181 keeps working); the all-reduce is either called
explicitly after backward (`allreduce_gradients`, used inside the captured training step) or
attached as post-accumulate-grad hooks (`attach_allreduce_hooks`) so that the unchanged
ModelWrapper.fit loop (wrapper.py:51-74) trains data-parallel.
BatchNorm1d statistics in GraphNet stay per replica (as torch DDP's default).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_sets(x: torch.Tensor, idx: torch.Tensor, y: torch.Tensor, rank: int, world: int):
    """Contiguous slice of whole sets for `rank` out of a ragged batch (x[sum N_i, F], idx, y[B, .]);
    idx is re-based to start at 0 (layout of utils/data.py:651-663)."""
    B = y.shape[0]
    per = (B + world - 1) // world
    b0, b1 = min(rank * per, B), min((rank + 1) * per, B)
    counts = torch.bincount(idx, minlength=B)
    off = torch.zeros(B + 1, dtype=torch.int64, device=idx.device)
    off[1:] = torch.cumsum(counts, 0)
    r0, r1 = int(off[b0]), int(off[b1])
    return x[r0:r1], idx[r0:r1] - b0, y[b0:b1]


def flatten_grads(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])


def unflatten_into_grads(flat: torch.Tensor, params: Sequence[torch.nn.Parameter]) -> None:
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int | None = None) -> None:
    """Average .grad over all ranks with ONE flat fp32 bucket (payload 0.3-0.8 MB for the yaml models:
    latency bound, so a single collective beats per-tensor calls)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = world or dist.get_world_size()
    if world == 1:
        return
    params = [p for p in params if p.requires_grad]
    flat = flatten_grads(params)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)
    unflatten_into_grads(flat, params)


def attach_allreduce_hooks(model: torch.nn.Module) -> List:
    """Data-parallel training under the UNCHANGED reference loop: when the last parameter gradient of
    a backward pass has been accumulated, all gradients are averaged across ranks (one flat bucket).
    Returns the hook handles."""
    params = [p for p in model.parameters() if p.requires_grad]
    state = {"seen": 0}

    def hook(_p):
        state["seen"] += 1
        if state["seen"] == len(params):
            state["seen"] = 0
            allreduce_gradients(params)

    return [p.register_post_accumulate_grad_hook(hook) for p in params]
