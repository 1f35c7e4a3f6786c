"""CPU, world_size 2, gloo: host-side logic of the data-parallel path (sharding of ragged
batches by whole sets, flat-bucket gradient averaging, hook-driven all-reduce).  The CUDA
kernels are not involved; NCCL runs the same code on the GPU box (bench.py --gpus N)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pcc_b200 import distributed as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    # a ragged batch of 5 sets; a tiny plain-torch "model" stands in for the CUDA modules
    sizes = [3, 1, 4, 2, 5]
    x = torch.randn(sum(sizes), 3)
    idx = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    y = torch.rand(len(sizes), 1)
    xs, ids, ys = D.shard_sets(x, idx, y, rank, world)
    lin = torch.nn.Linear(3, 1)
    with torch.no_grad():
        lin.weight.fill_(0.5); lin.bias.fill_(0.1)

    def local_loss(xa, ia, ya):
        n_sets = ya.shape[0]
        pooled = torch.zeros(n_sets, 3).index_add(0, ia, xa)
        return ((lin(pooled) - ya) ** 2).sum()   # sum: so that summed shard losses == full-batch loss

    # (1) explicit flat-bucket all-reduce: average of per-rank grads == full-batch grad / world
    lin.zero_grad()
    local_loss(xs, ids, ys).backward()
    D.allreduce_gradients(lin.parameters())
    got = [p.grad.clone() for p in lin.parameters()]
    lin.zero_grad()
    local_loss(x, idx, y).backward()
    ref = [p.grad / world for p in lin.parameters()]
    ok1 = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, ref))
    # (2) hook-driven all-reduce under an unchanged "loss.backward()" loop
    handles = D.attach_allreduce_hooks(lin)
    lin.zero_grad()
    local_loss(xs, ids, ys).backward()
    ok2 = all(torch.allclose(p.grad, b, atol=1e-6) for p, b in zip(lin.parameters(), ref))
    for h in handles:
        h.remove()
    # (2b) a parameter that gets no gradient in one pass (unused branch) must neither skip that pass's all-reduce
    #      nor shift the next one: two modules, the second one only used in the second pass
    both = torch.nn.ModuleList([lin, torch.nn.Linear(3, 1)])
    with torch.no_grad():
        both[1].weight.fill_(-0.25); both[1].bias.fill_(0.3)
    handles = D.attach_allreduce_hooks(both)
    both.zero_grad()
    local_loss(xs, ids, ys).backward()                       # pass 1: both[1] unused -> no grad for it
    ok2 = ok2 and all(torch.allclose(p.grad, b, atol=1e-6) for p, b in zip(lin.parameters(), ref))
    both.zero_grad()

    def loss2(xa, ia, ya):
        pooled = torch.zeros(ya.shape[0], 3).index_add(0, ia, xa)
        return ((lin(pooled) + both[1](pooled) - ya) ** 2).sum()
    loss2(xs, ids, ys).backward()                            # pass 2: every parameter used
    got2 = [p.grad.clone() for p in both.parameters()]
    for h in handles:
        h.remove()
    both.zero_grad()
    loss2(x, idx, y).backward()
    ok2 = ok2 and all(torch.allclose(a, p.grad / world, atol=1e-6) for a, p in zip(got2, both.parameters()))
    # (3) shards are disjoint, ordered, re-based
    total_rows = torch.tensor([xs.shape[0]])
    dist.all_reduce(total_rows)
    ok3 = int(total_rows) == x.shape[0] and (ids.numel() == 0 or int(ids.min()) == 0)
    q.put((rank, ok1, ok2, ok3))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_data_parallel_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, ok1, ok2, ok3 in res:
        assert ok1, f"rank {rank}: flat-bucket average != full-batch gradient / world"
        assert ok2, f"rank {rank}: hook-driven all-reduce mismatch"
        assert ok3, f"rank {rank}: shard_sets lost rows or did not re-base idx"
