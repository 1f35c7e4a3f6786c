"""Minimal train-step loop for ncu (dev tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
act, pool, res = (sys.argv[1], sys.argv[2], sys.argv[3] == "1") if len(sys.argv) > 3 else ("relu", "max", False)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
B, N, d, out = 256, 1024, 3, 10
torch.manual_seed(0)
m = pcc_b200.DeepSets(d, [256, 256], [256], out, act, layer_norm=False, residual_block=res, pooling=pool, precision="bf16").cuda()
x = torch.randn(B * N, d, device="cuda"); idx = torch.arange(B, device="cuda").repeat_interleave(N)
y = (torch.rand(B, out, device="cuda") > 0.5).float(); lossf = torch.nn.BCEWithLogitsLoss()
for _ in range(steps):
    loss = lossf(m(x, idx, num_sets=B), y); m.zero_grad(set_to_none=True); loss.backward()
torch.cuda.synchronize(); print("ok", float(loss))
