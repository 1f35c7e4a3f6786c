"""Warm per-kernel durations of one DeepSets train step (dev tool): torch.profiler (CUPTI) over eager
launches of the bench.py step, so every kernel of the step — not only the three big fused ones — gets a
device-side duration.  usage: python tools/kernel_times.py [act pool res] [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200.train_step import GraphedTrainStep

act, pool, res = (sys.argv[1], sys.argv[2], sys.argv[3] == "1") if len(sys.argv) > 3 else ("relu", "max", False)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
B, N, d, out = 256, 1024, 3, 10
torch.manual_seed(0)
m = pcc_b200.DeepSets(d, [256, 256], [256], out, act, layer_norm=False, residual_block=res, pooling=pool, precision="bf16").cuda()
x = torch.randn(B * N, d, device="cuda"); idx = torch.arange(B, device="cuda").repeat_interleave(N)
y = (torch.rand(B, out, device="cuda") > 0.5).float()
gs = GraphedTrainStep(m, (x, idx), y, forward_kwargs={"num_sets": B}, use_graph=False)
for _ in range(3):
    gs.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        gs.run()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t > 0:
        rows.append((t / steps, e.count / steps, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"# {act} {pool} res={int(res)}: sum of kernel time {tot:.1f} us/step, {sum(r[1] for r in rows):.0f} launches/step")
for t, c, k in rows:
    print(f"{t:9.2f} us  x{c:4.1f}  {k[:100]}")
# individual launches of the small tile kernels, in launch order (last step)
evs = [e for e in prof.events() if "head_tile" in e.name]
evs.sort(key=lambda e: e.time_range.start)
for e in evs[-4:]:
    print(f"   head_tile launch: {e.device_time if hasattr(e, 'device_time') else e.cuda_time:.2f} us")
