import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200 import _lib
pool = sys.argv[1] if len(sys.argv) > 1 else "sum"
act = sys.argv[2] if len(sys.argv) > 2 else "relu"
res = len(sys.argv) > 3 and sys.argv[3] == "res"
B, N = 256, 1024
m = pcc_b200.DeepSets(3, [256, 256], [256], 10, act, layer_norm=False, residual_block=res, pooling=pool, precision="bf16").cuda()
x = torch.randn(B * N, 3, device="cuda"); idx = torch.arange(B, device="cuda").repeat_interleave(N)
y = (torch.rand(B, 10, device="cuda") > 0.5).float(); lf = torch.nn.BCEWithLogitsLoss()
buf = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2: _lib.call("pcc_debug_set_trace", _lib.ptr(buf))
    loss = lf(m(x, idx, num_sets=B), y); m.zero_grad(set_to_none=True)
    if it == 2:
        buf.zero_()   # drop the forward kernel's events; keep the backward's
    loss.backward()
torch.cuda.synchronize(); _lib.call("pcc_debug_set_trace", None)
t = buf.cpu().numpy().reshape(4, 2048, 2)
names = {0: "tile start", 1: "acquire done", 2: "x staged+arrive", 3: "dZfinal built", 4: "dZfinal stored+arrive", 5: "accA(z0) ready",
         6: "h0 epi done", 7: "h0 stored+arrive", 8: "accB ready", 9: "accA(z1) ready", 10: "acquire done", 11: "fused pass done",
         12: "stores issued", 13: "acquire done", 14: "accB/accA ready", 15: "acquire done", 16: "dZ0 done", 17: "dZ0 stored"}
ev = [(int(t[0, i, 1]), int(t[0, i, 0])) for i in range(2048) if t[0, i, 1] != 0]
t0 = ev[0][0]; prev = None; tiles = 0
for ts, id_ in ev:
    if id_ == 0: tiles += 1
    if tiles == 2:
        print(f"{ts - t0:9d} (+{0 if prev is None else ts - prev:6d}) {names.get(id_, id_)}")
    prev = ts
print("tiles", tiles, "total", ev[-1][0] - t0)
