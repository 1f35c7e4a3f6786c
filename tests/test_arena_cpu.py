"""CPU: host logic of the gradient arena (pcc_b200/distributed.py) — slice layout, adoption of arena views as
.grad by autograd, and the flat bucket the all-reduce runs on.  The kernels that write into the arena on the
GPU are covered by the -m gpu tests and tools/test_peer_allreduce.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))

from pcc_b200.distributed import GradArena, grad_like  # noqa: E402


class _ArenaLinear(torch.autograd.Function):
    """stand-in for the package's autograd Functions: gradients are written into grad_like() buffers"""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w, b)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        dw, db = grad_like(w), grad_like(b)
        dw.copy_(dy.t() @ x)
        db.copy_(dy.sum(0))
        return None, dw, db


def test_arena_layout_and_adoption():
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)          # 15 + 3 parameters: the bias slice starts 16-byte aligned
    arena = GradArena(list(lin.parameters()))
    assert arena.numel % 4 == 0 and arena.numel >= 18
    offs = sorted(v[0] for v in arena.slices.values())
    assert offs == [0, 16]
    assert arena.view_for(torch.zeros(3, 5)) is None          # unknown tensor: not an arena slice
    x = torch.randn(7, 5)
    assert grad_like(lin.weight).data_ptr() != arena.flat.data_ptr()   # no arena active: fresh tensors
    with arena:
        _ArenaLinear.apply(x, lin.weight, lin.bias).sum().backward()
    assert arena.holds_all_grads()                                # autograd adopted the views as .grad
    ref = torch.nn.Linear(5, 3)
    ref.load_state_dict(lin.state_dict())
    ref(x).sum().backward()
    torch.testing.assert_close(lin.weight.grad, ref.weight.grad)
    torch.testing.assert_close(lin.bias.grad, ref.bias.grad)
    # the flat bucket IS the gradients: scaling it in place (what the all-reduce average does) scales .grad
    arena.flat.mul_(0.5)
    torch.testing.assert_close(lin.weight.grad, 0.5 * ref.weight.grad)
    # a gradient that did not come through the arena is detected
    lin.bias.grad = torch.zeros(3)
    assert not arena.holds_all_grads()
