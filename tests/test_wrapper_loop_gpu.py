"""GPU: the reference's training-loop sequence (models/wrapper.py:51-74 — .to(device), model(*inputs), zero_grad,
BCEWithLogitsLoss, backward, torch AdamW step, loss.item()) on the drop-in module, 20 steps over rotating batches,
against the same loop on the CPU oracle: the LOSS TRAJECTORY must match (fp32 mode tightly, bf16 mode within the
stated bf16 tolerance — errors compound over the optimizer steps)."""
import pytest
import torch

from helpers import ragged_batch
from oracle import deepsets_oracle as O

import pcc_b200

pytestmark = pytest.mark.gpu


def _oracle_trajectory(cfg, sd, batches, steps, lr):
    params = {k: v.clone() for k, v in sd.items()}
    names = list(params)
    leaves = [torch.nn.Parameter(params[k]) for k in names]
    opt = torch.optim.AdamW(leaves, lr=lr)
    losses = []
    for i in range(steps):
        x, idx, y = batches[i % len(batches)]
        cur = {k: p.detach() for k, p in zip(names, leaves)}
        _, loss, grads, _ = O.deepsets_train_step(cur, cfg, x, idx, y)
        opt.zero_grad()
        for k, p in zip(names, leaves):
            p.grad = grads[k]
        opt.step()
        losses.append(float(loss))
    return losses


@pytest.mark.parametrize("precision,act,pool,res,tol", [("fp32", "relu", "max", False, 2e-4), ("fp32", "gelu", "mean", True, 2e-4),
                                                        ("bf16", "gelu", "mean", True, 1e-2), ("bf16", "relu", "max", False, 3e-2)])
def test_wrapper_loop_loss_trajectory_matches_oracle(precision, act, pool, res, tol):
    d, H, out, steps, lr = 3, 128, 2, 20, 1e-3
    cfg = dict(input_dim=d, phi_layers=[H, H], rho_layers=[64], output_dim=out, activation=act, layer_norm=False,
               residual_block=res, pooling=pool)
    sd = O.init_state_dict(cfg, seed=5)
    batches = []
    for b in range(4):
        sizes = [200, 57, 128, 300, 64, 131]
        x, idx = ragged_batch(sizes, d, seed=40 + b)
        y = (torch.rand(len(sizes), out, generator=torch.Generator().manual_seed(50 + b)) > 0.5).float()
        batches.append((x, idx, y))
    ref = _oracle_trajectory(cfg, sd, batches, steps, lr)

    model = pcc_b200.DeepSets(**cfg, precision=precision).cuda()
    model.load_state_dict(sd)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=lr)            # wrapper.py:33
    criterion = torch.nn.BCEWithLogitsLoss()                       # wrapper.py:38
    got = []
    for i in range(steps):
        x, idx, y = batches[i % len(batches)]
        inputs = [t.to("cuda") for t in (x, idx) if t is not None]  # wrapper.py:54
        yd = y.to("cuda")                                          # wrapper.py:55
        logits = model(*inputs)                                    # wrapper.py:58
        opt.zero_grad()                                            # wrapper.py:61
        loss = criterion(logits, yd)                               # wrapper.py:64
        loss.backward()                                            # wrapper.py:67
        opt.step()                                                 # wrapper.py:70
        got.append(loss.item())                                    # wrapper.py:73
    worst = max(abs(a - b) / abs(b) for a, b in zip(got, ref))
    print(f"wrapper loop {precision} {act}/{pool}: loss {ref[0]:.5f} -> {ref[-1]:.5f} (oracle), worst relative deviation {worst:.2e}")
    assert worst < tol
