"""GPU parity for the graph neighbour stage: kNN vs oracle/knn_oracle.py (bit-exact
indices and distances), CSR, GraphConv aggregation and the full GraphNet train step vs
oracle/graphnet_oracle.py (PyG semantics restated; parity unpinned — see its header)."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import graphnet_oracle as GO
from oracle import knn_oracle as KO

import pcc_b200
from pcc_b200 import functional as PF

pytestmark = pytest.mark.gpu


def _clouds(sizes, seed, F=4):
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    feats = torch.randn(n, F, generator=g)
    feats[:, 0] = torch.rand(n, generator=g)
    memb = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return feats, memb, off


@pytest.mark.parametrize("sizes,k", [([64, 100, 33], 20), ([1024, 1024], 20), ([21, 22, 500], 20), ([5, 3, 1, 40], 4),
                                     ([300], 32), ([10, 15], 20), ([2500, 1500, 90], 20), ([5000, 700], 8)])
def test_knn_bit_exact(sizes, k):
    """identical neighbour ids and identical fp32 distance bits; clouds of several 1024-candidate tiles included"""
    feats, memb, off = _clouds(sizes, seed=7)
    ref_nbr, ref_d2 = KO.knn_neighbours(feats[:, 1:4].numpy(), off, k)
    nbr, d2 = PF.knn(feats.cuda()[:, 1:4], torch.from_numpy(off).cuda(), k)
    assert np.array_equal(nbr.cpu().numpy(), ref_nbr)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)  # identical fp32 bits (no FMA contraction)


def test_knn_ties_break_by_lower_id():
    """points on a coarse integer lattice (many exactly equal distances, repeated points): the (d2, id) order of the oracle
    (the kernel compares distance bits only and relies on the ascending scan order for the ids)"""
    g = torch.Generator().manual_seed(11)
    sizes = [400, 257]
    pos = torch.randint(0, 4, (sum(sizes), 3), generator=g).float()
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ref_nbr, ref_d2 = KO.knn_neighbours(pos.numpy(), off, 20)
    nbr, d2 = PF.knn(pos.cuda(), torch.from_numpy(off).cuda(), 20)
    assert np.array_equal(nbr.cpu().numpy(), ref_nbr)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)


def test_knn_graph_edge_convention():
    feats, memb, off = _clouds([50, 70], seed=8)
    ei, _ = pcc_b200.knn_graph(feats.cuda(), memb.cuda(), k=20)
    ref = KO.knn_edges(KO.knn_neighbours(feats[:, 1:4].numpy(), off, 20)[0])
    assert np.array_equal(ei.cpu().numpy(), ref)
    assert ei.shape == (2, 120 * 20)
    assert bool((ei[0] != ei[1]).all())  # no self loops


def test_csr_build_sorted_rows():
    g = torch.Generator().manual_seed(1)
    n, E = 500, 7000
    edges = torch.randint(0, n, (2, E), generator=g)
    csr = PF.GraphCSR(edges.cuda(), n)
    rowptr, perm = (t.cpu() for t in csr.by_dst)
    counts = torch.bincount(edges[1], minlength=n)
    assert torch.equal(rowptr[1:] - rowptr[:-1], counts)
    for i in (0, 17, 499):
        seg = perm[rowptr[i]:rowptr[i + 1]].long()
        assert torch.equal(seg, torch.nonzero(edges[1] == i).flatten())  # ascending edge ids


@pytest.mark.parametrize("aggr", ["add", "mean", "max"])
@pytest.mark.parametrize("C,use_w", [(4, False), (128, True), (1, True), (6, False)])
def test_graph_aggregate_fwd_bwd(aggr, C, use_w):
    g = torch.Generator().manual_seed(2)
    n, E = 300, 4000
    edges = torch.randint(0, n, (2, E), generator=g)
    edges[1, edges[1] == 5] = 6  # node 5 has no incoming edge -> aggregates to 0
    x = torch.randn(n, C, generator=g)
    w = torch.rand(E, generator=g) if use_w else None
    xr = x.clone().requires_grad_(True)
    ref = GO.graph_aggregate(xr, edges, w, aggr)
    go = torch.randn(n, C, generator=g)
    ref.backward(go)
    xc = x.cuda().requires_grad_(True)
    out = PF.graph_aggregate(xc, w.cuda() if use_w else None, PF.GraphCSR(edges.cuda(), n), aggr)
    out.backward(go.cuda())
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    assert rel_err(xc.grad, xr.grad) < 1e-5


@pytest.mark.parametrize("act,aggr,deepchem,use_w,hidden,F", [
    ("tanh", "add", True, False, 128, 4),   # configs/graph_net.yaml
    ("relu", "mean", False, True, 64, 4),
    ("gelu", "max", True, True, 64, 1),
    ("tanh", "add", False, False, 256, 4),
])
def test_graphnet_train_step_matches_oracle(act, aggr, deepchem, use_w, hidden, F):
    cfg = dict(input_dim=F, hidden_dim=hidden, output_dim=1, activation=act, use_gat=False, gat_heads=4,
               sag_pool=False, pool_ratio=0.5, local_pooling=aggr, global_pooling="mean", deepchem_style=deepchem)
    sizes = [60, 45, 80, 33]
    feats, memb, off = _clouds(sizes, seed=21, F=max(F, 4))
    nbr, _ = KO.knn_neighbours(feats[:, 1:4].numpy(), off, 8)
    edges = torch.from_numpy(KO.knn_edges(nbr))
    x = feats[:, :F].contiguous()
    gen = torch.Generator().manual_seed(22)
    w = torch.rand(edges.shape[1], generator=gen) if use_w else None
    y = (torch.rand(len(sizes), 1, generator=gen) > 0.5).float()
    sd = GO.init_state_dict(cfg, seed=23)
    ref_logits, ref_loss, ref_grads, ref_stats = GO.graphnet_train_step(sd, cfg, x, memb, edges, w, y)

    m = pcc_b200.GraphNet(**cfg).cuda()
    m.load_state_dict(sd)
    m.train()
    args = [x.cuda(), memb.cuda(), edges.cuda()] + ([w.cuda()] if use_w else [])
    logits = m(*args)
    loss = torch.nn.BCEWithLogitsLoss()(logits, y.cuda())
    loss.backward()
    torch.testing.assert_close(logits.detach().cpu(), ref_logits, rtol=1e-4, atol=1e-5)
    for k, ref in ref_grads.items():
        got = dict(m.named_parameters())[k].grad
        assert got is not None and rel_err(got, ref) < 2e-4, (k, rel_err(got, ref))
    new_sd = m.state_dict()
    for k, v in ref_stats.items():
        torch.testing.assert_close(new_sd[k].cpu(), v, rtol=1e-4, atol=1e-6)
    # eval mode uses the running statistics
    m.eval()
    with torch.no_grad():
        ev = m(*args)
    sd_eval = {k: v.cpu() for k, v in m.state_dict().items()}
    ref_ev = GO.graphnet_forward(sd_eval, cfg, x, memb, edges, w, training=False)
    torch.testing.assert_close(ev.cpu(), ref_ev, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["graphnet_yaml_tanh_add_deepchem", "graphnet_relu_mean_weights",
                                  "graphnet_gelu_max_weights_deepchem"])
def test_graphnet_matches_reference_module_golden(name):
    """CUDA GraphNet (fp32 mode) against outputs of the reference's own models/graph_net.py (run with a loop-based
    stand-in for the absent torch_geometric kernels, oracle/gen_golden_graphnet.py): logits, every gradient,
    BatchNorm running statistics, eval logits; arbitrary edge order, ragged in-degrees, isolated nodes."""
    from helpers import load_graphnet_golden
    g = load_graphnet_golden(name)
    m = pcc_b200.GraphNet(**g["cfg"]).cuda()
    assert set(m.state_dict()) == set(g["sd"])
    m.load_state_dict(g["sd"])
    m.train()
    args = [g["x"].cuda(), g["membership"].cuda(), g["edges"].cuda()] + ([g["weights"].cuda()] if g["weights"] is not None else [])
    logits = m(*args)
    torch.nn.BCEWithLogitsLoss()(logits, g["y"].cuda()).backward()
    torch.testing.assert_close(logits.detach().cpu(), g["logits"], rtol=1e-4, atol=1e-5)
    for k, ref in g["grads"].items():
        got = dict(m.named_parameters())[k].grad.cpu()
        assert rel_err(got, ref) < 2e-4 or float((got - ref).abs().max()) < 2e-7, (k, rel_err(got, ref))
    sd = m.state_dict()
    for k, ref in g["after"].items():
        torch.testing.assert_close(sd[k].cpu(), ref, rtol=1e-4, atol=1e-6)
    m.eval()
    with torch.no_grad():
        torch.testing.assert_close(m(*args).cpu(), g["logits_eval"], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_graphnet_tf32_dense_mode():
    """functional.set_dense_precision("tf32"): single-TF32 tensor-core GEMMs for the node-level layers.  Stated
    tolerance: operands rounded to 10 mantissa bits -> logits within 2e-2 of max|ref|, gradients within 5e-2
    (relative Frobenius); the default 3xTF32 mode is the one held to 1e-4 above."""
    from helpers import rel_l2
    cfg = dict(input_dim=4, hidden_dim=128, output_dim=1, activation="tanh", use_gat=False, gat_heads=4,
               sag_pool=False, pool_ratio=0.5, local_pooling="add", global_pooling="mean", deepchem_style=True)
    sizes = [300, 200, 400, 256]          # > 128 nodes: the large-tile kernels are the ones that switch
    feats, memb, off = _clouds(sizes, seed=31, F=4)
    nbr, _ = KO.knn_neighbours(feats[:, 1:4].numpy(), off, 8)
    edges = torch.from_numpy(KO.knn_edges(nbr))
    gen = torch.Generator().manual_seed(32)
    y = (torch.rand(len(sizes), 1, generator=gen) > 0.5).float()
    sd = GO.init_state_dict(cfg, seed=33)
    ref_logits, _, ref_grads, _ = GO.graphnet_train_step(sd, cfg, feats, memb, edges, None, y)
    m = pcc_b200.GraphNet(**cfg).cuda()
    m.load_state_dict(sd)
    m.train()
    PF.set_dense_precision("tf32")
    try:
        logits = m(feats.cuda(), memb.cuda(), edges.cuda())
        torch.nn.BCEWithLogitsLoss()(logits, y.cuda()).backward()
        torch.cuda.synchronize()
    finally:
        PF.set_dense_precision("fp32")
    assert rel_err(logits.detach().cpu(), ref_logits) < 2e-2
    for k, ref in ref_grads.items():
        assert rel_l2(dict(m.named_parameters())[k].grad, ref) < 5e-2, k


@pytest.mark.gpu
def test_gaussian_edge_weights_match_reference_golden_and_oracle():
    """pcc_edge_weights (utils/data.py:835-845 on device, batched): edge lengths and the per-graph median sigma are
    bit-exact against the oracle (which is pinned to the reference's own function by tests/golden/edge_weights.npz);
    the weights differ only by the exp implementation (a few ulp: rtol 2e-6)."""
    import os
    from oracle import edge_weights_oracle as EO
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edge_weights.npz"))
    feats, edges, eoff, gold, node0 = [], [], [0], [], 0
    for gi in range(int(z["count"])):
        f, e = z[f"features_{gi}"], z[f"edges_{gi}"]
        feats.append(f); edges.append(e + node0); gold.append(z[f"weights_{gi}"])
        node0 += f.shape[0]; eoff.append(eoff[-1] + e.shape[1])
    feats, edges, gold = np.concatenate(feats), np.concatenate(edges, axis=1), np.concatenate(gold)
    eoff = np.asarray(eoff, dtype=np.int64)
    ref_w, ref_sigma, ref_d = EO.compute_weights_batched(feats, edges, eoff)
    assert np.array_equal(ref_w, gold)                      # batched oracle == reference goldens
    f_gpu = torch.from_numpy(feats).cuda()
    w, sigma = PF.edge_weights(f_gpu[:, 1:4], torch.from_numpy(edges).cuda(), torch.from_numpy(eoff).cuda(),
                               return_sigma=True)
    assert np.array_equal(sigma.cpu().numpy(), ref_sigma)   # exact medians (odd and even counts)
    np.testing.assert_allclose(w.cpu().numpy(), gold, rtol=2e-6, atol=1e-30)
    # through the module-level helper on a kNN batch (membership -> per-graph edge ranges), incl. an empty graph id
    sizes = [300, 77, 512]
    fb, memb, off = _clouds(sizes, seed=41, F=4)
    nbr, _ = KO.knn_neighbours(fb[:, 1:4].numpy(), off, 8)
    e2 = KO.knn_edges(nbr)
    eo2 = np.asarray([0] + list(np.cumsum([s * 8 for s in sizes])), dtype=np.int64)
    w2_ref, _, _ = EO.compute_weights_batched(fb.numpy(), e2, eo2)
    w2 = pcc_b200.gaussian_edge_weights(fb.cuda(), torch.from_numpy(e2).cuda(), memb.cuda())
    np.testing.assert_allclose(w2.cpu().numpy(), w2_ref, rtol=2e-6, atol=1e-30)
