"""GPU parity: CUDA DeepSets (through the C-ABI library) vs the reference goldens and
the CPU oracle.  fp32 path: logits and every parameter gradient within rtol 1e-4
(north_star); max-pool argmax rows bit-exact on these tie-free inputs."""
import itertools

import pytest
import torch

from helpers import golden_cases, load_golden, ragged_batch, rel_err
from oracle import deepsets_oracle as O

import pcc_b200
from pcc_b200 import functional as PF

pytestmark = pytest.mark.gpu
RTOL_FP32 = 1e-4


def _run_cuda(cfg, sd, x, idx, y, precision="fp32"):
    m = pcc_b200.DeepSets(**cfg, precision=precision).cuda()
    m.load_state_dict(sd)
    logits = m(x.cuda(), idx.cuda())
    loss = torch.nn.BCEWithLogitsLoss()(logits, y.cuda())
    m.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    return m, logits.detach().cpu(), float(loss), grads


@pytest.mark.parametrize("name", golden_cases())
def test_fp32_matches_reference_golden(name):
    g = load_golden(name)
    m, logits, loss, grads = _run_cuda(g["cfg"], g["sd"], g["x"], g["idx"], g["y"])
    assert m.last_path == "fp32"
    torch.testing.assert_close(logits, g["logits"], rtol=RTOL_FP32, atol=1e-5)
    assert abs(loss - g["loss"]) < 1e-5
    for k, ref in g["grads"].items():
        assert grads[k] is not None and rel_err(grads[k], ref) < RTOL_FP32, (k, rel_err(grads[k], ref))


@pytest.mark.parametrize("name", [n for n in golden_cases() if "max" in n])
def test_argmax_bit_exact(name):
    g = load_golden(name)
    m = pcc_b200.DeepSets(**g["cfg"], precision="fp32").cuda()
    m.load_state_dict(g["sd"])
    x, idx = g["x"].cuda(), g["idx"].cuda()
    phi_x = m._mlp(m._phi_plan, x)
    off = PF.segment_offsets(idx, int(g["idx"].max()) + 1)
    pooled, arg = PF.segment_pool(phi_x, off, "max", return_argmax=True)
    assert torch.equal(arg.cpu().long(), g["argmax"])


@pytest.mark.parametrize("act,pool,ln,res", list(itertools.product(["relu", "gelu", "silu"], ["sum", "mean", "max"],
                                                                   [False, True], [False, True])))
def test_fp32_matches_oracle_all_variants(act, pool, ln, res):
    cfg = dict(input_dim=5, phi_layers=[48, 48, 32], rho_layers=[24], output_dim=3, activation=act, layer_norm=ln,
               residual_block=res, pooling=pool)
    sd = O.init_state_dict(cfg, seed=11)
    x, idx = ragged_batch([1, 31, 32, 33, 127, 128, 129, 300], 5, seed=12)
    y = (torch.rand(8, 3, generator=torch.Generator().manual_seed(13)) > 0.5).float()
    ref_logits, ref_loss, ref_grads, _ = O.deepsets_train_step(sd, cfg, x, idx, y)
    _, logits, loss, grads = _run_cuda(cfg, sd, x, idx, y)
    torch.testing.assert_close(logits, ref_logits, rtol=RTOL_FP32, atol=1e-5)
    for k, ref in ref_grads.items():
        assert rel_err(grads[k], ref) < RTOL_FP32, (k, rel_err(grads[k], ref))


def test_segment_offsets_and_index_max():
    idx = torch.tensor([0, 0, 0, 1, 3, 3, 3, 3, 5], dtype=torch.long).cuda()
    assert PF.index_max(idx) == 5
    off = PF.segment_offsets(idx, 6).cpu()
    assert off.tolist() == [0, 3, 4, 4, 8, 8, 9]
    # unsorted idx: only the histogram matters (deep_sets.py:91-92)
    perm = torch.randperm(9, generator=torch.Generator().manual_seed(0)).cuda()
    assert PF.segment_offsets(idx[perm], 6).cpu().tolist() == off.tolist()
    big = torch.arange(3000).repeat_interleave(7).cuda()
    assert torch.equal(PF.segment_offsets(big, 3000).cpu(), torch.arange(3001) * 7)


def test_permutation_invariance_within_sets():
    cfg = dict(input_dim=3, phi_layers=[32, 32], rho_layers=[16], output_dim=2, activation="gelu", layer_norm=False,
               residual_block=True, pooling="max")
    sd = O.init_state_dict(cfg, seed=3)
    sizes = [40, 77, 128]
    x, idx = ragged_batch(sizes, 3, seed=4)
    m = pcc_b200.DeepSets(**cfg, precision="fp32").cuda()
    m.load_state_dict(sd)
    a = m(x.cuda(), idx.cuda())
    xs, s = [], 0
    for i, n in enumerate(sizes):
        p = torch.randperm(n, generator=torch.Generator().manual_seed(50 + i))
        xs.append(x[s:s + n][p])
        s += n
    b = m(torch.cat(xs).cuda(), idx.cuda())
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)  # max pool: exact up to the rho GEMM order


def test_large_equal_sets_properties():
    """BASELINE config-2 shape (B=256,N=1024,H=256): size-independent checks — pooled max
    equals a direct torch max of the CUDA phi output, argmax rows lie inside their set."""
    cfg = dict(input_dim=3, phi_layers=[256, 256], rho_layers=[256], output_dim=10, activation="relu",
               layer_norm=False, residual_block=False, pooling="max")
    m = pcc_b200.DeepSets(**cfg, precision="fp32").cuda()
    B, N = 256, 1024
    x = torch.randn(B * N, 3, device="cuda")
    idx = torch.arange(B, device="cuda").repeat_interleave(N)
    phi_x = m._mlp(m._phi_plan, x)
    off = PF.segment_offsets(idx, B)
    pooled, arg = PF.segment_pool(phi_x, off, "max", return_argmax=True)
    ref, ref_arg = phi_x.view(B, N, -1).max(dim=1)
    assert torch.equal(pooled, ref)
    lo = (torch.arange(B, device="cuda") * N).view(B, 1)
    assert bool(((arg >= lo) & (arg < lo + N)).all())
    assert torch.equal(phi_x[arg.long(), torch.arange(256, device="cuda").expand(B, -1)], pooled)


def test_fused_bce_loss_and_gather():
    g = torch.Generator().manual_seed(5)
    z = (torch.randn(256, 10, generator=g) * 3).cuda().requires_grad_(True)
    y = (torch.rand(256, 10, generator=g) > 0.5).float().cuda()
    loss = PF.bce_with_logits(z, y)
    (loss * 2.0).backward()
    z2 = z.detach().clone().requires_grad_(True)
    ref = torch.nn.BCEWithLogitsLoss()(z2, y)
    (ref * 2.0).backward()
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(z.grad, z2.grad, rtol=1e-5, atol=1e-8)
    zb = (torch.randn(2500, 10, generator=g) * 3).cuda()     # > 4 K logits: the multi-block (atomic) variant
    yb = (torch.rand(2500, 10, generator=g) > 0.5).float().cuda()
    torch.testing.assert_close(PF.bce_with_logits(zb, yb), torch.nn.BCEWithLogitsLoss()(zb, yb), rtol=1e-5, atol=1e-6)
    x = torch.randn(1000, 3, generator=g).cuda()
    idx = torch.randint(-1, 1000, (777,), generator=g).int().cuda()
    assert torch.equal(PF.gather_rows(x, idx), x[idx.long().clamp_min(0)])


@pytest.mark.parametrize("dims,act,M", [([256, 256, 10], "relu", 256), ([128, 64, 32, 1], "gelu", 37),
                                        ([256, 1], "silu", 5), ([1024, 512, 128, 256, 3], "tanh", 70),
                                        ([10, 7, 3], "relu", 33), ([256, 256, 10], "gelu", 0)])
def test_fused_head_matches_torch(dims, act, M):
    g = torch.Generator().manual_seed(9)
    ws, ps = [], []
    for i in range(len(dims) - 1):
        w = (torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).cuda().requires_grad_(True)
        b = (torch.randn(dims[i + 1], generator=g) * 0.1).cuda().requires_grad_(True)
        ws += [w, b]
    x = torch.randn(M, dims[0], generator=g).cuda().requires_grad_(True)
    go = torch.randn(M, dims[-1], generator=g).cuda()
    assert PF.head_supported(dims, act)
    y = PF.mlp_head(x, act, ws)
    y.backward(go)
    got = [x.grad.clone()] + [t.grad.clone() for t in ws]
    x.grad = None
    for t in ws:
        t.grad = None
    fn = {"relu": torch.relu, "gelu": torch.nn.functional.gelu, "silu": torch.nn.functional.silu, "tanh": torch.tanh}[act]
    h = x
    for i in range(len(dims) - 1):
        h = torch.nn.functional.linear(h, ws[2 * i], ws[2 * i + 1])
        if i < len(dims) - 2:
            h = fn(h)
    h.backward(go)
    torch.testing.assert_close(y, h, rtol=1e-4, atol=1e-5)
    for a, b in zip(got, [x.grad] + [t.grad for t in ws]):
        assert rel_err(a, b) < 1e-4
