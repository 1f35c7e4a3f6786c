"""TEST INFRASTRUCTURE ONLY — CPU oracle of the optimizer step of the reference's training loop.

The reference constructs `torch.optim.Adam(params, lr=lr)` or `torch.optim.AdamW(params, lr=lr)` with default
hyper-parameters (/root/reference/models/wrapper.py:30-33) and calls `optimizer.step()` once per batch (:70).  The
algorithm lives in the reference's dependency torch (2.11 here; torch/optim/adam.py `_single_tensor_adam`, adamw.py),
restated below in float32 numpy, operation for operation:
    AdamW:  p <- p * (1 - lr * wd)                      Adam:  g <- g + wd * p
    m <- m + (g - m) * (1 - beta1)                      (exp_avg.lerp_)
    v <- v * beta2 + (1 - beta2) * g * g                (exp_avg_sq.mul_().addcmul_())
    bc1 = 1 - beta1^t,  bc2 = 1 - beta2^t               (Python floats = float64)
    p <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
PINNED: tests/test_oracle_golden.py runs this restatement next to torch.optim.Adam / AdamW themselves (the dependency
is installed in this image, so the real implementation is the fixture) on seeded parameters and gradients; the GPU
test (tests/test_optim_gpu.py) compares the CUDA kernel with torch.optim directly."""
import math

import numpy as np


class AdamOracle:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=None, decoupled=True):
        self.p = [np.array(p, dtype=np.float32) for p in params]
        self.m = [np.zeros_like(p) for p in self.p]
        self.v = [np.zeros_like(p) for p in self.p]
        self.t = [0 for _ in self.p]   # torch keeps one step count per parameter
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.wd = (1e-2 if decoupled else 0.0) if weight_decay is None else weight_decay
        self.decoupled = decoupled

    def step(self, grads):
        f = np.float32
        for i, g in enumerate(grads):
            if g is None:
                continue
            g = np.array(g, dtype=np.float32)
            self.t[i] += 1
            p, m, v = self.p[i], self.m[i], self.v[i]
            if self.wd != 0.0:
                if self.decoupled:
                    p *= f(1.0 - self.lr * self.wd)
                else:
                    g = g + f(self.wd) * p
            m += (g - m) * f(1.0 - self.b1)
            v *= f(self.b2)
            v += f(1.0 - self.b2) * g * g
            bc1 = 1.0 - self.b1 ** self.t[i]
            bc2_sqrt = math.sqrt(1.0 - self.b2 ** self.t[i])
            denom = np.sqrt(v) / f(bc2_sqrt) + f(self.eps)
            p -= f(self.lr / bc1) * (m / denom)
        return self.p
