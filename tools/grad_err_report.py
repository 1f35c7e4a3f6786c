"""Per-tensor bf16-path errors against the fp32 oracle (diagnostic; feeds the tolerances in tests/test_fused_gpu.py).

For every fused parity case and for the two full-size configurations (B=256, N=1024: relu+max, yaml gelu+res+mean)
prints rel-L2 error of every parameter gradient and the max-norm error of the logits.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "point-cloud-classifier_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import pcc_b200  # noqa: E402
from helpers import ragged_batch, rel_err, rel_l2  # noqa: E402
from oracle import deepsets_oracle as O  # noqa: E402
from pcc_b200 import functional as PF  # noqa: E402
from test_fused_gpu import CASES, _cfg, _fused_argmax  # noqa: E402


def one(cfg, sizes, d, seed, out=3, tag=""):
    sd = O.init_state_dict(cfg, seed=seed)
    x, idx = ragged_batch(sizes, d, seed=seed + 1)
    y = (torch.rand(len(sizes), out, generator=torch.Generator().manual_seed(seed + 2)) > 0.5).float()
    m = pcc_b200.DeepSets(**cfg, precision="bf16").cuda()
    m.load_state_dict(sd)
    arg = None
    if cfg["pooling"] == "max":
        off = PF.segment_offsets(idx.cuda(), len(sizes))
        arg = _fused_argmax(m, x.cuda(), off, cfg["activation"])
    ref_logits, _, ref_grads, _ = O.deepsets_train_step(sd, cfg, x, idx, y, argmax_rows=arg)
    q_logits, _, free_grads, _ = O.deepsets_train_step(sd, cfg, x, idx, y, phi_operand_rounding="bf16", argmax_rows=arg)
    logits = m(x.cuda(), idx.cuda())
    torch.nn.BCEWithLogitsLoss()(logits, y.cuda()).backward()
    row = [f"{tag:34s} logits {rel_err(logits, ref_logits):.1e}/{rel_err(logits, q_logits):.1e}"]
    for k, ref in ref_grads.items():
        got = dict(m.named_parameters())[k].grad
        s = f"{k}={rel_l2(got, ref):.1e}"
        if free_grads is not None:
            s += f"/{rel_l2(got, free_grads[k]):.1e}"
        row.append(s)
    print("  ".join(row), flush=True)


def main():
    for act, pool, res, H, depth, d, sizes in CASES:
        one(_cfg(act, pool, res, H, depth, d), sizes, d, 51, tag=f"{act}/{pool}/res={int(res)}/H{H}/L{depth}/d{d}/B{len(sizes)}")
    big = [1024] * 256
    one(dict(input_dim=3, phi_layers=[256, 256], rho_layers=[256], output_dim=10, activation="relu", layer_norm=False,
             residual_block=False, pooling="max"), big, 3, 71, out=10, tag="FULL relu/max B256 N1024")
    one(dict(input_dim=6, phi_layers=[256, 256], rho_layers=[256], output_dim=1, activation="gelu", layer_norm=False,
             residual_block=True, pooling="mean"), big, 6, 73, out=1, tag="FULL yaml gelu/res/mean B256 N1024")


if __name__ == "__main__":
    main()
