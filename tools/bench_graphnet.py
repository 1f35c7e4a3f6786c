"""Secondary measurement (BASELINE config 4): GraphNet (configs/graph_net.yaml model) on kNN graphs,
k=20, N=1024 points per cloud; kNN build and train step timed separately with CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200
from pcc_b200 import functional as PF


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


def run(B, N=1024, k=20):
    torch.manual_seed(0)
    m = pcc_b200.GraphNet(input_dim=4, hidden_dim=128, output_dim=1, activation="tanh", local_pooling="add",
                          global_pooling="mean", deepchem_style=True).cuda()
    n = B * N
    feats = torch.randn(n, 4, device="cuda")
    feats[:, 0] = torch.rand(n, device="cuda")
    memb = torch.arange(B, device="cuda").repeat_interleave(N)
    y = (torch.rand(B, 1, device="cuda") > 0.5).float()
    off = PF.segment_offsets(memb, B)
    t_knn = timeit(lambda: PF.knn(feats[:, 1:4], off, k))
    nbr, _ = PF.knn(feats[:, 1:4], off, k)
    edges = PF.knn_edges(nbr)
    lf = torch.nn.BCEWithLogitsLoss()

    def step():
        loss = lf(m(feats, memb, edges, num_graphs=B), y)
        m.zero_grad(set_to_none=True)
        loss.backward()
    t_step = timeit(step, steps=5, warmup=2)
    # the same step captured in a CUDA graph (101 launches per step: eager mode is launch bound)
    t_graph = float("nan")
    try:
        from pcc_b200.train_step import GraphedTrainStep
        gs = GraphedTrainStep(m, [feats, memb, edges], y, forward_kwargs={"num_graphs": B})
        t_graph = timeit(gs.run, steps=5, warmup=2)
    except Exception as e:  # noqa: BLE001
        print(f"  (graph capture failed: {type(e).__name__}: {e})")
    pairs = B * N * N
    print(f"GraphNet B={B:4d} N={N} k={k}: kNN {t_knn:8.3f} ms ({pairs / t_knn / 1e6:8.1f} Gpairs/s) | "
          f"train step eager {t_step:8.3f} ms, CUDA graph {t_graph:8.3f} ms = {B / t_graph * 1e3:8.0f} graphs/s, "
          f"E={edges.shape[1]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        PF.set_dense_precision(sys.argv[1])   # "tf32": single-TF32 node-level GEMMs
        print(f"dense precision: {sys.argv[1]}")
    run(32)
    run(256)
    for N in (256, 4096):
        run(262144 // N, N=N)
