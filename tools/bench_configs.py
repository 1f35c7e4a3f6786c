"""Secondary measurements (not the driver's bench): CUDA-graph step time and per-kernel times for the other
BASELINE configs — yaml default (gelu + residual + mean), sum pooling, ragged log-normal set sizes (C3),
fp32 parity path.  Prints one line per configuration."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200 import _lib
from pcc_b200.train_step import GraphedTrainStep


def sizes_for(kind, B, N, seed=2):
    if kind == "equal":
        return [N] * B
    g = torch.Generator().manual_seed(seed)  # C3: N_i = clamp(round(exp(N(ln 1500, 0.8^2))), 16, 4096)
    s = torch.exp(torch.randn(B, generator=g) * 0.8 + torch.log(torch.tensor(1500.0))).round().clamp(16, 4096)
    return [int(v) for v in s]


def run(name, act, pool, res, precision, d=3, out=10, B=256, N=1024, kind="equal", steps=30):
    torch.manual_seed(0)
    m = pcc_b200.DeepSets(d, [256, 256], [256], out, act, layer_norm=False, residual_block=res, pooling=pool,
                          precision=precision).cuda()
    sizes = sizes_for(kind, B, N)
    n = sum(sizes)
    x = torch.randn(n, d, device="cuda")
    idx = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)]).cuda()
    y = (torch.rand(B, out, device="cuda") > 0.5).float()
    gs = GraphedTrainStep(m, [x, idx], y, forward_kwargs={"num_sets": B})
    for _ in range(5):
        gs.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gs.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eager = GraphedTrainStep(m, [x, idx], y, forward_kwargs={"num_sets": B}, use_graph=False, warmup=2)
    _lib.call("pcc_prof_enable", 1)
    for _ in range(10):
        eager.run()
    torch.cuda.synchronize()
    _lib.call("pcc_prof_enable", 0)
    k = []
    for slot in range(3):
        t, c = C.c_double(0), C.c_int64(0)
        _lib.call("pcc_prof_read", slot, C.byref(t), C.byref(c))
        k.append(round(t.value / c.value, 4) if c.value else None)
    print(f"{name:34s} {precision:4s} path={m.last_path:10s} points={n:7d} step {ms:7.3f} ms = {B / ms * 1e3:9.0f} samples/s "
          f"= {n / ms / 1e3:7.1f} Mpts/s | kernels fwd/chain/wgrad ms = {k}", flush=True)


if __name__ == "__main__":
    run("C2 relu+max (headline)", "relu", "max", False, "bf16")
    run("C2' yaml gelu+res+mean d=6 out=1", "gelu", "mean", True, "bf16", d=6, out=1)
    run("C2'' silu+sum", "silu", "sum", False, "bf16")
    run("C3 ragged lognormal, sum", "relu", "sum", False, "bf16", kind="ragged")
    run("C3 ragged lognormal, mean gelu", "gelu", "mean", True, "bf16", kind="ragged")
    run("C3 ragged lognormal, max", "relu", "max", False, "bf16", kind="ragged")
    run("C1 shape B=32 yaml", "gelu", "mean", True, "bf16", d=6, out=1, B=32)
    run("C2 relu+max fp32 parity path", "relu", "max", False, "fp32", steps=5)
