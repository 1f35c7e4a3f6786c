"""kNN build (pcc_knn, k = 20): Gpairs/s over the point-count sweep and the fraction of the MEASURED FP32 FMA peak
(8 flop per pair — 3 sub, 3 mul, 2 add — against the FMA pipe measured on this box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
from pcc_b200 import _lib, functional as PF


def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


blocks, iters = 148 * 8, 20000
out = torch.empty(blocks * 256, device="cuda")
ms = timeit(lambda: _lib.call("pcc_selftest_fp32_peak", _lib.ptr(out), blocks, iters, 0, _lib.stream_ptr(0)), steps=5)
peak = blocks * 256 * 16.0 * iters / (ms * 1e-3) / 1e12
print(f"measured FP32 FMA peak: {peak:.1f} TFLOP/s ({ms:.3f} ms for {blocks} x 256 threads x {iters} x 8 FMA)")
for N in (256, 1024, 4096, 16384):
    B = 262144 // N
    pos = torch.randn(B * N, 3, device="cuda")
    off = torch.arange(B + 1, device="cuda", dtype=torch.int64) * N
    t = timeit(lambda: PF.knn(pos, off, 20))
    pairs = B * N * N
    gp = pairs / t / 1e6
    print(f"kNN k=20 B={B:5d} N={N:6d}: {t:8.3f} ms = {gp:8.1f} Gpairs/s = {gp * 8 / 1e3:6.2f} TFLOP/s (8 flop/pair) = "
          f"{gp * 8 / 1e3 / peak * 100:5.1f} % of the measured FP32 peak")
