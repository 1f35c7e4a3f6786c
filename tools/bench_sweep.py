"""BASELINE config 5: point-count sweep N = 256 .. 16384 at constant points per step (B = 262144 / N), DeepSets
C2 model (phi [3-256-256]+final, rho [256]-10), bf16 fused path, CUDA-graph train step on one GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200.train_step import GraphedTrainStep


def run(N, pool, act="relu", total=262144, steps=40):
    B = total // N
    torch.manual_seed(0)
    m = pcc_b200.DeepSets(3, [256, 256], [256], 10, act, layer_norm=False, residual_block=False, pooling=pool,
                          precision="bf16").cuda()
    x = torch.randn(B * N, 3, device="cuda")
    idx = torch.arange(B, device="cuda").repeat_interleave(N)
    y = (torch.rand(B, 10, device="cuda") > 0.5).float()
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
    print(f"N={N:6d} B={B:5d} {act}+{pool:4s}: step {ms:7.3f} ms = {B / ms * 1e3:10.0f} samples/s = "
          f"{total / ms / 1e3:8.1f} Mpts/s  [{m.last_path}]", flush=True)


if __name__ == "__main__":
    for pool in ("max", "sum"):
        for N in (256, 512, 1024, 2048, 4096, 8192, 16384):
            run(N, pool)
