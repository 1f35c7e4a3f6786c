"""per-kernel time of one GraphNet train step (bf16 fused path, kNN inside), torch profiler over eager launches"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from torch.profiler import profile, ProfilerActivity

B, N, k = 256, 1024, 20
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
torch.manual_seed(0)
m = pcc_b200.KnnGraphNet(k=k, precision=precision, input_dim=4, hidden_dim=128, output_dim=1, activation="tanh", local_pooling="add",
                         global_pooling="mean", deepchem_style=True).cuda()
n = B * N
feats = torch.randn(n, 4, device="cuda"); feats[:, 0] = torch.rand(n, device="cuda")
memb = torch.arange(B, device="cuda").repeat_interleave(N)
y = (torch.rand(B, 1, device="cuda") > 0.5).float()
lf = torch.nn.BCEWithLogitsLoss()
def step():
    loss = lf(m(feats, memb, num_graphs=B), y)
    m.zero_grad(set_to_none=True)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows) / 3
print(f"# KnnGraphNet {precision} B={B} N={N} k={k}: {tot:.0f} us/step of kernels, path={m.net.last_path}")
for e in rows[:32]:
    print(f"{e.device_time_total / 3:9.1f} us x{e.count / 3:4.1f}  {e.key[:110]}")
