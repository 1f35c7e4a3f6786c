"""Generate tests/golden/graphnet_*.npz by running the UNMODIFIED reference module
/root/reference/models/graph_net.py (constructor + forward control flow, :10-104) on CPU.

Run in the build container only:   python oracle/gen_golden_graphnet.py

The reference imports `torch_geometric.nn` (graph_net.py:5), which is not vendored, not version-pinned and not
installable here.  This script registers a STAND-IN module for it whose GraphConv / global_*_pool follow PyG's
published semantics and parameter names, written as plain per-node python loops (deliberately NOT the vectorised
index_add / scatter formulation of oracle/graphnet_oracle.py, so that the two restatements check each other):

  GraphConv(in, out, aggr): lin_rel = Linear(in, out, bias=True), lin_root = Linear(in, out, bias=False);
      out_i = lin_rel(aggr_{e: edge_index[1,e] = i} w_e * x[edge_index[0,e]]) + lin_root(x_i); aggr add / mean / max;
      a node without incoming edges aggregates to 0.
  global_mean_pool(x, batch): per-graph mean of rows, size = batch.max() + 1.

What the goldens therefore pin: everything that IS in the reference — the control flow of graph_net.py (activation
before BatchNorm :75-76, deepchem ordering :86-100, the hard-coded global_mean_pool :92,:96, fc1 width 256 :61),
torch's own BatchNorm1d / Linear / activations, the state_dict key set.  The third-party kernels stay "parity
unpinned" (stated in oracle/graphnet_oracle.py).  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import types

sys.dont_write_bytecode = True
import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("PCC_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _LoopGraphConv(nn.Module):
    def __init__(self, in_channels, out_channels, aggr="add"):
        super().__init__()
        self.aggr = aggr
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index, edge_weight=None):
        n = x.shape[0]
        incoming = [[] for _ in range(n)]
        for e in range(edge_index.shape[1]):
            incoming[int(edge_index[1, e])].append(e)
        rows = []
        for i in range(n):
            if not incoming[i]:
                rows.append(x.new_zeros(x.shape[1]))
                continue
            es = torch.tensor(incoming[i], dtype=torch.long)
            msg = x[edge_index[0, es]]
            if edge_weight is not None:
                msg = msg * edge_weight[es].view(-1, 1)
            rows.append(msg.sum(0) if self.aggr == "add" else (msg.mean(0) if self.aggr == "mean" else msg.max(0)[0]))
        return self.lin_rel(torch.stack(rows)) + self.lin_root(x)


def _global_pool(kind):
    def pool(x, batch):
        B = int(batch.max()) + 1
        out = []
        for b in range(B):
            rows = x[batch == b]
            out.append(rows.mean(0) if kind == "mean" else (rows.sum(0) if kind == "add" else rows.max(0)[0]))
        return torch.stack(out)
    return pool


def _unavailable(*a, **k):
    raise NotImplementedError("GATConv / SAGPooling are not restated")


def install_stub():
    tg, tgnn = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn")
    tgnn.GraphConv, tgnn.GATConv, tgnn.SAGPooling = _LoopGraphConv, _unavailable, _unavailable
    tgnn.global_mean_pool, tgnn.global_add_pool, tgnn.global_max_pool = (_global_pool(k) for k in ("mean", "add", "max"))
    tg.nn = tgnn
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = tg, tgnn


CASES = [
    # name, cfg, nodes per graph, in-degree k, edge weights?, seed
    ("yaml_tanh_add_deepchem", dict(input_dim=4, hidden_dim=128, output_dim=1, activation="tanh", use_gat=False, gat_heads=4,
                                    sag_pool=False, pool_ratio=0.5, local_pooling="add", global_pooling="mean",
                                    deepchem_style=True), [40, 25, 33], 5, False, 1),
    ("relu_mean_weights", dict(input_dim=4, hidden_dim=64, output_dim=2, activation="relu", local_pooling="mean",
                               deepchem_style=False), [30, 30, 17, 9], 4, True, 2),
    ("gelu_max_weights_deepchem", dict(input_dim=1, hidden_dim=64, output_dim=1, activation="gelu", local_pooling="max",
                                       global_pooling="add", deepchem_style=True), [21, 50], 6, True, 3),
]


def random_graph_batch(sizes, k, F_in, with_w, g):
    xs, es, ms, off = [], [], [], 0
    for gi, n in enumerate(sizes):
        xs.append(torch.randn(n, F_in, generator=g))
        dst = torch.arange(n).repeat_interleave(k)
        src = torch.randint(0, n, (n * k,), generator=g)
        keep = torch.rand(n * k, generator=g) > 0.15          # ragged in-degrees, some isolated nodes
        keep &= dst != 0                                      # node 0 of every graph: no incoming edge
        es.append(torch.stack([src[keep], dst[keep]]) + off)
        ms.append(torch.full((n,), gi, dtype=torch.long))
        off += n
    edges = torch.cat(es, dim=1)
    perm = torch.randperm(edges.shape[1], generator=g)        # edge order is arbitrary in the reference's collate too
    edges = edges[:, perm].contiguous()
    w = torch.rand(edges.shape[1], generator=g) if with_w else None
    return torch.cat(xs), torch.cat(ms), edges, w


def main():
    install_stub()
    sys.path.insert(0, REF)
    from models.graph_net import GraphNet  # the reference module, unmodified

    os.makedirs(OUT, exist_ok=True)
    for name, cfg, sizes, k, with_w, seed in CASES:
        torch.manual_seed(seed)
        model = GraphNet(**cfg)
        model.train()
        g = torch.Generator().manual_seed(200 + seed)
        x, mem, edges, w = random_graph_batch(sizes, k, cfg["input_dim"], with_w, g)
        y = (torch.rand(len(sizes), cfg["output_dim"], generator=g) > 0.5).float()
        sd0 = {kk: v.detach().clone() for kk, v in model.state_dict().items()}
        args = (x, mem, edges) + ((w,) if with_w else ())
        logits = model(*args)
        loss = torch.nn.BCEWithLogitsLoss()(logits, y)
        model.zero_grad()
        loss.backward()
        out = {"cfg_json": np.array(json.dumps(cfg)), "x": x.numpy(), "membership": mem.numpy(), "edges": edges.numpy(),
               "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.array(float(loss))}
        if with_w:
            out["weights"] = w.numpy()
        for kk, v in sd0.items():
            out["sd/" + kk] = v.numpy()
        for kk, p in model.named_parameters():
            out["grad/" + kk] = p.grad.numpy()
        for kk, v in model.state_dict().items():               # BatchNorm running statistics after the step
            if "running" in kk or "num_batches" in kk:
                out["after/" + kk] = v.numpy()
        model.eval()
        with torch.no_grad():
            out["logits_eval"] = model(*args).numpy()
        np.savez_compressed(os.path.join(OUT, f"graphnet_{name}.npz"), **out)
        print(f"graphnet_{name}: nodes {x.shape[0]} edges {edges.shape[1]} loss {float(loss):.6f} keys {len(sd0)}")


if __name__ == "__main__":
    main()
