"""Quick device timing of one DeepSets train step (dev tool; bench.py is the contract)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
import pcc_b200

def run(act, pool, res, precision, B=256, N=1024, d=3, out=10, steps=20):
    torch.manual_seed(0)
    m = pcc_b200.DeepSets(d, [256, 256], [256], out, act, layer_norm=False, residual_block=res, pooling=pool,
                          precision=precision).cuda()
    x = torch.randn(B * N, d, device="cuda")
    idx = torch.arange(B, device="cuda").repeat_interleave(N)
    y = (torch.rand(B, out, device="cuda") > 0.5).float()
    lossf = torch.nn.BCEWithLogitsLoss()
    def step():
        logits = m(x, idx, num_sets=B)
        loss = lossf(logits, y)
        m.zero_grad(set_to_none=True)
        loss.backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # forward only
    with torch.no_grad():
        for _ in range(3): m(x, idx, num_sets=B)
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps): m(x, idx, num_sets=B)
        e1.record(); torch.cuda.synchronize()
    fms = e0.elapsed_time(e1) / steps
    print(f"{act:5s} {pool:5s} res={int(res)} {precision}: train {ms:8.3f} ms/step = {B/ms*1e3:10.0f} samples/s | fwd {fms:7.3f} ms  [{m.last_path}]", flush=True)

if __name__ == "__main__":
    for prec in ("bf16", "fp32"):
        run("relu", "max", False, prec)
        run("gelu", "mean", True, prec)
