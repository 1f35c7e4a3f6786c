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


class GradArena:
    """ONE flat fp32 buffer that holds every parameter gradient of a model.

    While an arena is active (`with arena:`), the backward kernels of this package write dW / db straight
    into its slices (the autograd Functions ask `grad_like`), and autograd adopts those views as `.grad`
    (AccumulateGrad steals a fresh contiguous view when `.grad` is None).  The data-parallel all-reduce then
    runs in place on `flat` — no flatten / unflatten copies, no extra launches."""

    _active: "GradArena | None" = None

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.slices = {}
        off = 0
        for p in self.params:
            off = (off + 3) // 4 * 4            # 16-byte aligned slices
            self.slices[p.data_ptr()] = (off, p.numel(), tuple(p.shape))
            off += p.numel()
        self.numel = (off + 3) // 4 * 4
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def view_for(self, t: torch.Tensor):
        ent = self.slices.get(t.data_ptr())
        if ent is None or t.dtype != torch.float32 or tuple(t.shape) != ent[2]:
            return None
        return self.flat.narrow(0, ent[0], ent[1]).view(ent[2])  # a fresh view each time (so that it can be adopted)

    def holds_all_grads(self) -> bool:
        for p in self.params:
            ent = self.slices[p.data_ptr()]
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * ent[0]:
                return False
        return True

    def __enter__(self):
        GradArena._active = self
        return self

    def __exit__(self, *exc):
        GradArena._active = None


def grad_like(t: torch.Tensor) -> torch.Tensor:
    """gradient buffer for parameter tensor `t`: its slice of the active GradArena, else a fresh tensor"""
    a = GradArena._active
    if a is not None:
        v = a.view_for(t)
        if v is not None:
            return v
    return torch.empty_like(t)


class PeerAllReduce:
    """Average a flat fp32 bucket over the ranks of one node with the library's own one-shot all-reduce over
    NVLink peer memory (pcc_peer.cu).  A plain kernel launch: it can be captured into the CUDA graph of the
    train step.  The process group is only used once, to exchange the CUDA IPC handles."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import ctypes as C
        from . import _lib as L
        self.L, self.C = L, C
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.numel = (numel + 3) // 4 * 4
        self.buf_bytes = (self.numel * 4 + 255) // 256 * 256
        self.local, self.opened = None, []
        # Every rank runs the same collective sequence (gather handles, gather status, barrier) whatever fails
        # locally, so a rank that cannot allocate or map IPC memory makes ALL ranks raise together instead of
        # leaving the healthy ones in a barrier.
        err = None
        handle = (C.c_ubyte * 64)()
        try:
            region = C.c_void_p()
            L.call("pcc_peer_alloc", self.buf_bytes, C.byref(region), C.cast(handle, C.c_void_p), self.dev_index)
            self.local = region.value
        except Exception as e:
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if err is None else None, group=group)
        self.regions = (C.c_void_p * self.world)()
        if err is None and all(h is not None for h in handles):
            try:
                for r in range(self.world):
                    if r == self.rank:
                        self.regions[r] = self.local
                    else:
                        peer = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                        L.call("pcc_peer_open", C.cast(hb, C.c_void_p), C.byref(peer), self.dev_index)
                        self.regions[r] = peer.value
                        self.opened.append(peer.value)
            except Exception as e:
                err = e
        elif err is None:
            err = RuntimeError("a peer rank could not allocate its IPC region")
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            self.close()
            raise RuntimeError(f"peer all-reduce setup failed on rank(s) {[r for r, o in enumerate(oks) if not o]}"
                               + (f": {type(err).__name__}: {err}" if err is not None else ""))
        self.counters = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def run(self, flat: torch.Tensor) -> torch.Tensor:
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() == self.numel
        L, C = self.L, self.C
        L.call("pcc_peer_allreduce", L.ptr(flat), flat.numel(), C.cast(self.regions, C.c_void_p), self.buf_bytes,
               self.rank, self.world, 1.0 / self.world, L.ptr(self.counters), self.dev_index, L.stream_ptr(self.dev_index))
        return flat

    def close(self):
        if getattr(self, "local", None) is None and not getattr(self, "opened", None):
            return
        torch.cuda.synchronize(self.device)
        for ptr_ in self.opened:
            self.L.call("pcc_peer_close", self.C.c_void_p(ptr_), self.dev_index)
        if self.local is not None:
            self.L.call("pcc_peer_free", self.C.c_void_p(self.local), self.dev_index)
        self.local, self.opened = None, []


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
    """Data-parallel training under the UNCHANGED reference loop (wrapper.py:51-74): at the END of every backward
    pass all gradients are averaged across ranks (one flat bucket).  Returns the hook handles.

    The first parameter hook of a pass queues an end-of-backward callback on the autograd engine, so the
    collective fires exactly once per pass whatever subset of the parameters received a gradient (a parameter
    without one contributes zeros, as in `flatten_grads`): every rank issues the same collective sequence even
    when a branch leaves some parameters unused, and a backward that raises midway leaves no stale count
    behind (the flag is cleared when the next pass starts queuing)."""
    from torch.autograd import Variable
    params = [p for p in model.parameters() if p.requires_grad]
    state = {"queued": False}

    def finish():
        state["queued"] = False
        allreduce_gradients(params)

    def hook(_p):
        if not state["queued"]:
            state["queued"] = True
            try:
                Variable._execution_engine.queue_callback(finish)
            except Exception:
                state["queued"] = False
                raise

    return [p.register_post_accumulate_grad_hook(hook) for p in params]
