"""Shared loaders for the parity tests (test infrastructure)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix="deepsets_"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg_json"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    out = {"cfg": cfg, "sd": sd, "grads": grads, "x": torch.from_numpy(z["x"]), "idx": torch.from_numpy(z["idx"]),
           "y": torch.from_numpy(z["y"]), "logits": torch.from_numpy(z["logits"]), "loss": float(z["loss"])}
    if "argmax" in z.files:
        out["argmax"] = torch.from_numpy(z["argmax"])
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| — scale-aware error for gradient tensors."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0 if a.shape == b.shape else float("inf")
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def ragged_batch(sizes, d, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    x = torch.randn(n, d, generator=g)
    idx = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    return x.to(device), idx.to(device)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_F / ||b||_F — robust to the isolated flips (argmax / relu mask) that a change of
    operand precision causes."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load_graphnet_golden(name):
    """tests/golden/graphnet_*.npz: outputs of the reference's own models/graph_net.py (oracle/gen_golden_graphnet.py)"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {"cfg": json.loads(str(z["cfg_json"])),
           "sd": {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")},
           "grads": {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")},
           "after": {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")},
           "weights": torch.from_numpy(z["weights"]) if "weights" in z.files else None}
    for k in ("x", "membership", "edges", "y", "logits", "logits_eval"):
        out[k] = torch.from_numpy(z[k])
    out["loss"] = float(z["loss"])
    return out
