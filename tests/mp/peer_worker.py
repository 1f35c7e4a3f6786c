"""Worker of tests/test_multirank_gpu.py (one process per GPU, launched with torch.distributed.run):
(1) the library's peer-memory all-reduce against NCCL on random buckets, eager and inside a CUDA graph, results
    bitwise identical on every rank;
(2) a data-parallel GraphedTrainStep (gradient arena + peer all-reduce inside the captured step) against the
    average of the per-shard gradients computed without any collective.
Prints one line 'OK ...' on rank 0; any failed check raises (non-zero exit)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200  # noqa: E402
from pcc_b200.distributed import PeerAllReduce  # noqa: E402
from pcc_b200.train_step import GraphedTrainStep  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()


def same_on_all_ranks(t):
    chk = t.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    return all(float(c) == float(allc[0]) for c in allc)


# ---- (1) peer all-reduce vs NCCL
n = 199_428   # not a multiple of 4 on purpose -> padded
peer = PeerAllReduce(n, dev)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
worst = 0.0
for it in range(10):
    a = torch.randn(peer.numel, device=dev, generator=g)
    ref = a.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    peer.run(a)
    torch.cuda.synchronize()
    worst = max(worst, float((a - ref).abs().max()))
    assert same_on_all_ranks(a), "peer all-reduce result differs between ranks"
assert worst <= (0.0 if world == 2 else 1e-6), worst   # two ranks: one addition, same bits as NCCL
buf = torch.randn(peer.numel, device=dev, generator=g)
src = buf.clone()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gr = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        buf.copy_(src)
        peer.run(buf)
torch.cuda.synchronize()
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
ref = src.clone()
dist.all_reduce(ref)
ref /= world
assert float((buf - ref).abs().max()) <= 1e-6
peer.close()

# ---- (2) data-parallel captured train step vs per-shard gradients
Bl, N, d = 32, 256, 3
cfg = dict(input_dim=d, phi_layers=[256, 256], rho_layers=[256], output_dim=4, activation="relu", layer_norm=False,
           residual_block=False, pooling="max")
torch.manual_seed(7)
model = pcc_b200.DeepSets(**cfg, precision="bf16").to(dev)        # same seed: same parameters on every rank
gen = torch.Generator().manual_seed(11)
xs = [torch.randn(Bl * N, d, generator=gen).to(dev) for _ in range(world)]
ys = [(torch.rand(Bl, 4, generator=gen) > 0.5).float().to(dev) for _ in range(world)]
idx = torch.arange(Bl, device=dev).repeat_interleave(N)
per_shard = []
lf = torch.nn.BCEWithLogitsLoss()
for r in range(world):                                             # every rank computes every shard locally
    model.zero_grad(set_to_none=True)
    lf(model(xs[r], idx, num_sets=Bl), ys[r]).backward()
    per_shard.append([p.grad.detach().clone() for p in model.parameters()])
model.zero_grad(set_to_none=True)
gs = GraphedTrainStep(model, (xs[rank], idx), ys[rank], forward_kwargs={"num_sets": Bl}, allreduce=True)
assert gs.graph is not None and gs.peer is not None, "peer all-reduce inside the captured step expected"
gs.step((xs[rank], idx), ys[rank])
torch.cuda.synchronize()
worst = 0.0
for i, p in enumerate(model.parameters()):
    ref = sum(per_shard[r][i] for r in range(world)) / world
    err = float((p.grad - ref).norm() / (ref.norm() + 1e-30))
    worst = max(worst, err)
    assert err < 1e-5, (i, err)
assert same_on_all_ranks(gs.arena.flat), "averaged gradients differ between ranks"
dist.barrier()
if rank == 0:
    print(f"OK world {world}: peer vs nccl max diff {worst:.2e}; dp gradient == mean of shard gradients", flush=True)
gs.peer.close()
dist.destroy_process_group()
