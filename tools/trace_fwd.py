import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200 import _lib, functional as PF, fused as FZ
B, N = 256, 1024
CFG = sys.argv[1] if len(sys.argv) > 1 else "deepsets"      # deepsets (relu + max, d = 3) | yaml (gelu + res + mean, d = 6)
if CFG == "yaml":
    D, ACT_, POOL_ = 6, "gelu", "mean"
    m = pcc_b200.DeepSets(6, [256, 256], [256], 1, "gelu", layer_norm=False, residual_block=True, pooling="mean", precision="bf16").cuda()
else:
    D, ACT_, POOL_ = 3, "relu", "max"
    m = pcc_b200.DeepSets(3, [256, 256], [256], 10, "relu", layer_norm=False, pooling="max", precision="bf16").cuda()
x = torch.randn(B * N, D, device="cuda"); idx = torch.arange(B, device="cuda").repeat_interleave(N)
off = PF.segment_offsets(idx, B)
buf = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(2): FZ.phi_pool(x, off, m._phi_plan, ACT_, POOL_)
    _lib.call("pcc_debug_set_trace", _lib.ptr(buf))
    FZ.phi_pool(x, off, m._phi_plan, ACT_, POOL_)
    torch.cuda.synchronize()
    _lib.call("pcc_debug_set_trace", None)
t = buf.cpu().numpy().reshape(3, 2048, 2)
ev = []
for role in (0, 1, 2):
    for i in range(2048):
        if t[role, i, 1] == 0: break
        ev.append((int(t[role, i, 1]), role, int(t[role, i, 0])))
ev.sort()
t0 = ev[0][0]
# print tiles 5..6 worth of events
names = {0: "H image free, h0 start", 5: "H h0 done", 11: "H acc1 ready", 21: "H epi1 done", 30: "P accF ready", 40: "P pool done",
         101: "M L1 begin", 121: "M L1 first slab", 141: "M L1 issued", 102: "M L2 begin (pool drained)", 122: "M L2 first slab", 142: "M L2 issued"}
for l in range(3):
    names[100 + l] = f"M wait operand L{l}"; names[110 + l] = f"M operand L{l} ready"; names[120 + l] = f"M first slab L{l}"
    names[130 + l] = f"M last slab L{l}"; names[140 + l] = f"M issued L{l}"
cnt = 0
prev = None
for ts, role, id_ in ev:
    if id_ == 0: cnt += 1
    if 6 <= cnt <= 7:
        print(f"{ts - t0:9d} (+{0 if prev is None else ts - prev:6d})  {names.get(id_, id_)}")
        prev = ts
print("total cycles", ev[-1][0] - t0, "tiles", cnt)
