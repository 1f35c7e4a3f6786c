"""Generate tests/golden/deepsets_*.npz by running the UNMODIFIED reference module.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python oracle/gen_golden.py

For every case: build reference `DeepSets(**cfg)` (models/deep_sets.py:5-146) under
torch.manual_seed(seed), run forward + BCEWithLogitsLoss (models/wrapper.py:38) +
backward on CPU fp32, and store cfg, state_dict, inputs, logits, loss, every parameter
gradient and (for max pooling) the argmax rows recovered from the reference's own phi
output.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

REF = os.environ.get("PCC_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CASES = [
    # name, cfg, set sizes, seed
    ("relu_max", dict(input_dim=3, phi_layers=[64, 64], rho_layers=[64], output_dim=10, activation="relu",
                      layer_norm=False, residual_block=False, pooling="max"), [128] * 6, 1),
    ("yaml_gelu_res_mean", dict(input_dim=6, phi_layers=[64, 64], rho_layers=[64], output_dim=1, activation="gelu",
                                layer_norm=False, residual_block=True, sparse_batching=True, pooling="mean"),
     [33, 1, 200, 128, 129, 64, 7], 2),
    ("silu_ln_sum", dict(input_dim=4, phi_layers=[32, 48, 48], rho_layers=[32, 16], output_dim=1,
                         activation="silu", layer_norm=True, residual_block=True, pooling="sum"),
     [31, 32, 33, 127, 128, 129, 1, 300], 3),
    ("gelu_ln_max", dict(input_dim=3, phi_layers=[32], rho_layers=[32], output_dim=2, activation="gelu",
                         layer_norm=True, residual_block=False, pooling="max"), [5, 250, 64, 17], 4),
    ("relu_res_sum_wide", dict(input_dim=3, phi_layers=[256, 256], rho_layers=[256], output_dim=10,
                               activation="relu", layer_norm=False, residual_block=True, pooling="sum"),
     [130, 126], 5),
    ("relu_max_256", dict(input_dim=3, phi_layers=[256, 256], rho_layers=[256], output_dim=10,
                          activation="relu", layer_norm=False, residual_block=False, pooling="max"),
     [256, 100, 156], 6),
]


def main():
    sys.path.insert(0, REF)
    from models.deep_sets import DeepSets  # the reference module, unmodified

    os.makedirs(OUT, exist_ok=True)
    for name, cfg, sizes, seed in CASES:
        torch.manual_seed(seed)
        model = DeepSets(**cfg)
        g = torch.Generator().manual_seed(100 + seed)
        n = sum(sizes)
        x = torch.randn(n, cfg["input_dim"], generator=g)
        idx = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
        y = (torch.rand(len(sizes), cfg["output_dim"], generator=g) > 0.5).float()
        logits = model(x, idx)
        loss = torch.nn.BCEWithLogitsLoss()(logits, y)
        model.zero_grad()
        loss.backward()
        rec = {"cfg_json": np.array(json.dumps(cfg)), "x": x.numpy(), "idx": idx.numpy(), "y": y.numpy(),
               "logits": logits.detach().numpy(), "loss": loss.detach().numpy()}
        for k, v in model.state_dict().items():
            rec["sd/" + k] = v.numpy()
        for k, p in model.named_parameters():
            rec["grad/" + k] = p.grad.numpy()
        if cfg["pooling"] == "max":
            with torch.no_grad():
                phi_x = model.phi(x)
            off = np.concatenate([[0], np.cumsum(sizes)])
            arg = np.stack([phi_x[off[b]:off[b + 1]].max(dim=0)[1].numpy() + off[b] for b in range(len(sizes))])
            rec["argmax"] = arg.astype(np.int64)
        path = os.path.join(OUT, f"deepsets_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


if __name__ == "__main__":
    main()
