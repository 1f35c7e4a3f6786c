"""Warm per-kernel durations of one GraphNet train step (dev tool, torch.profiler)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200 import functional as PF
from torch.profiler import profile, ProfilerActivity
B, N, k = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 1024, 20
deep = (sys.argv[2] == "1") if len(sys.argv) > 2 else False
torch.manual_seed(0)
m = pcc_b200.GraphNet(input_dim=4, hidden_dim=128, output_dim=1, activation="tanh", local_pooling="add",
                      global_pooling="mean", deepchem_style=deep).cuda()
n = B * N
feats = torch.randn(n, 4, device="cuda"); feats[:, 0] = torch.rand(n, device="cuda")
memb = torch.arange(B, device="cuda").repeat_interleave(N)
y = (torch.rand(B, 1, device="cuda") > 0.5).float()
off = PF.segment_offsets(memb, B)
nbr, _ = PF.knn(feats[:, 1:4], off, k)
edges = PF.knn_edges(nbr)
lf = torch.nn.BCEWithLogitsLoss()
def step():
    loss = lf(m(feats, memb, edges, num_graphs=B), y); m.zero_grad(set_to_none=True); loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps): step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)
    if t > 0: rows.append((t / steps, e.count / steps, e.key))
rows.sort(reverse=True)
print(f"# GraphNet B={B} N={N} k={k} deepchem={deep}: {sum(r[0] for r in rows):.0f} us/step, {sum(r[1] for r in rows):.0f} launches")
for t, c, kk in rows[:28]: print(f"{t:9.1f} us x{c:4.1f}  {kk[:110]}")
