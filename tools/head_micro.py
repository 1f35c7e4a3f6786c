"""Kernel duration of the head tile kernel vs the contraction length (dev tool)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
from pcc_b200 import functional as PF
from torch.profiler import profile, ProfilerActivity
M = 256
for K in (4, 32, 128, 256, 512, 1024):
    for N in (32, 256):
        w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); x = torch.randn(M, K, device="cuda")
        for _ in range(3): PF.mlp_head(x, "relu", [w, b])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(20): PF.mlp_head(x, "relu", [w, b])
            torch.cuda.synchronize()
        t = [e for e in prof.key_averages() if "head_tile" in e.key][0]
        print(f"M={M} K={K:5d} N={N:4d}: {t.device_time_total / t.count:7.2f} us  ({M*N//1024} CTAs)")
