"""torchrun --nproc-per-node N tools/test_peer_allreduce.py : the library's peer-memory all-reduce against
NCCL on random buckets, eager and inside a CUDA graph, plus its device time."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
from pcc_b200.distributed import PeerAllReduce

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 199_428  # not a multiple of 4 on purpose -> padded
peer = PeerAllReduce(n, dev)
npad = peer.numel
g = torch.Generator(device=dev).manual_seed(1234 + rank)
worst = 0.0
for it in range(20):
    a = torch.randn(npad, device=dev, generator=g)
    ref = a.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    peer.run(a)
    torch.cuda.synchronize()
    worst = max(worst, float((a - ref).abs().max()))
# identical bits on every rank
chk = a.double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(c) == float(allc[0]) for c in allc)
# inside a CUDA graph
buf = torch.randn(npad, device=dev, generator=g)
src = buf.clone()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gr = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        buf.copy_(src)
        peer.run(buf)
torch.cuda.synchronize()
gr.replay(); torch.cuda.synchronize()
ref = src.clone(); dist.all_reduce(ref); ref /= world
gerr = float((buf - ref).abs().max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(200):
    gr.replay()
e1.record(); torch.cuda.synchronize()
t_peer = e0.elapsed_time(e1) / 200 * 1e3
e0.record()
for _ in range(200):
    dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) / 200 * 1e3
if rank == 0:
    print(f"world {world}: max|peer - nccl| eager {worst:.3e}, graph {gerr:.3e}, identical on all ranks: {same}; "
          f"copy+peer all-reduce (graph) {t_peer:.1f} us, NCCL all-reduce (eager) {t_nccl:.1f} us, {npad * 4 / 1e6:.2f} MB")
peer.close()
dist.destroy_process_group()
