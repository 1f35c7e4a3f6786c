import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcc_b200
from oracle import deepsets_oracle as O
from helpers import ragged_batch, rel_err, rel_l2

def one(act, pool, res, H, depth, d, sizes):
    cfg = dict(input_dim=d, phi_layers=[H] * depth, rho_layers=[64], output_dim=3, activation=act, layer_norm=False,
               residual_block=res, pooling=pool)
    sd = O.init_state_dict(cfg, seed=51)
    x, idx = ragged_batch(sizes, d, seed=52)
    y = (torch.rand(len(sizes), 3, generator=torch.Generator().manual_seed(53)) > 0.5).float()
    ref_logits, ref_loss, ref_grads, _ = O.deepsets_train_step(sd, cfg, x, idx, y)
    if pool == "max":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import test_fused_gpu as T
        from pcc_b200 import functional as PF
        m0 = pcc_b200.DeepSets(**cfg, precision="bf16").cuda(); m0.load_state_dict(sd)
        arg = T._fused_argmax(m0, x.cuda(), PF.segment_offsets(idx.cuda(), len(sizes)), act)
        ref_logits, ref_grads = T._oracle_step_with_argmax(sd, cfg, x, idx, y, arg)
    out = {}
    for prec in ("bf16", "fp32"):
        m = pcc_b200.DeepSets(**cfg, precision=prec).cuda(); m.load_state_dict(sd)
        logits = m(x.cuda(), idx.cuda())
        loss = torch.nn.BCEWithLogitsLoss()(logits, y.cuda()); loss.backward()
        torch.cuda.synchronize()
        out[prec] = {k: rel_l2(p.grad, ref_grads[k]) for k, p in m.named_parameters()}
        out[prec]["logits"] = rel_err(logits, ref_logits)
    print(f"--- {act}/{pool}/res={res}/H={H}/depth={depth}/d={d}/sets={len(sizes)} n={sum(sizes)}")
    for k in out["bf16"]:
        print(f"   {k:22s} bf16 {out['bf16'][k]:.2e}   fp32 {out['fp32'][k]:.2e}")
    sys.stdout.flush()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        one("relu", "max", False, 256, 2, 3, [1024, 1024, 1024])
        one("relu", "sum", False, 256, 1, 3, [256, 100, 156])
        one("gelu", "mean", True, 256, 2, 6, [33, 1, 200, 128, 129, 64, 7, 500])
        one("silu", "sum", True, 128, 2, 4, [31, 32, 33, 127, 128, 129, 1, 300])
        one("relu", "mean", False, 128, 2, 3, [128])
    elif which == "mid":
        one(sys.argv[3], sys.argv[4], False, int(sys.argv[2]), 2, 3, [1024] * int(sys.argv[5]))
    else:
        one("relu", "mean", False, 128, 2, 3, [1024] * int(which))
