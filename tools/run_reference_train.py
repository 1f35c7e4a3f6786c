"""Run the reference's own train.py unchanged on top of the B200 modules (integration recipe,
SURVEY.md §0/§7).  Needs the reference checkout (default /root/reference) — build container only.

  python tools/run_reference_train.py --model deep_sets --workdir /tmp/pcc_run --epochs 1
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_missing(names):
    for n in names:
        try:
            __import__(n)
        except Exception:
            m = types.ModuleType(n)
            if n == "torch_geometric":
                nn = types.ModuleType("torch_geometric.nn")
                for k in ("GraphConv", "GATConv", "SAGPooling", "global_mean_pool", "global_add_pool", "global_max_pool"):
                    setattr(nn, k, object)
                m.nn = nn
                sys.modules["torch_geometric.nn"] = nn
            if n == "matplotlib":
                sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
            sys.modules[n] = m


def write_synthetic_s2ppc(data_dir, events=24, seed=0):
    """npz schema of utils/data.py:599-609 / :633-641"""
    rng = np.random.default_rng(seed)
    for split in ("train", "val", "test"):
        d = os.path.join(data_dir, "S2PPC", split)
        os.makedirs(d, exist_ok=True)
        ev, rows = [], []
        for e in range(events):
            n = int(rng.integers(40, 200))
            ev += [e] * n
            rows.append(rng.normal(size=(n, 6)))
        rows = np.concatenate(rows)
        ev = np.array(ev)
        label = (ev % 2).astype(np.int64)
        np.savez(os.path.join(d, f"S2PPC_{split}_1.npz"), event_id=ev, energy=np.abs(rows[:, 0]) + 0.02,
                 energy_total=np.abs(rows[:, 1]) + 1.0, position_x=rows[:, 2], position_y=rows[:, 3],
                 position_z=rows[:, 4], time=np.abs(rows[:, 5]), label=label)


def write_synthetic_s2pg(data_dir, graphs=24, seed=0, k=6):
    """per-graph npz schema of utils/data.py:1112-1121 (read back at :1183-1196): features [n,4] fp32 (col 0 energy,
    cols 1:4 xyz), edges [2,E] int64 (source; target), weights [E] fp32, label, event_id"""
    rng = np.random.default_rng(seed)
    for split in ("train", "val", "test"):
        d = os.path.join(data_dir, "S2PG", split)
        os.makedirs(d, exist_ok=True)
        for i in range(graphs):
            n = int(rng.integers(30, 120))
            f = rng.normal(size=(n, 4)).astype(np.float32)
            f[:, 0] = rng.random(n).astype(np.float32)
            d2 = ((f[:, None, 1:4] - f[None, :, 1:4]) ** 2).sum(-1)
            np.fill_diagonal(d2, np.inf)
            nbr = np.argsort(d2, axis=1)[:, :k]
            edges = np.stack([nbr.reshape(-1), np.repeat(np.arange(n), k)]).astype(np.int64)
            dist = np.sqrt(d2[edges[1], edges[0]]).astype(np.float32)
            w = np.exp(-dist ** 2 / (2 * (np.median(dist) + 1e-6) ** 2)).astype(np.float32)
            np.savez(os.path.join(d, f"graph_{i:05d}.npz"), features=f, edges=edges, weights=w,
                     label=np.float32(i % 2), event_id=np.int64(i))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("PCC_REFERENCE", "/root/reference"))
    ap.add_argument("--model", default="deep_sets")
    ap.add_argument("--dataset", default="s2ppc")
    ap.add_argument("--workdir", default="/tmp/pcc_run")
    ap.add_argument("--epochs", type=int, default=1)
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    sys.path[:0] = [os.path.join(ROOT, "point-cloud-classifier_b200"), args.reference]
    stub_missing(["h5py", "matplotlib", "seaborn", "torch_geometric"])
    os.makedirs(args.workdir, exist_ok=True)
    if not os.path.exists(os.path.join(args.workdir, "configs")):
        os.symlink(os.path.join(args.reference, "configs"), os.path.join(args.workdir, "configs"))
    write_synthetic_s2ppc(os.path.join(args.workdir, "data", "continuous"))
    write_synthetic_s2pg(os.path.join(args.workdir, "data", "continuous"))
    os.chdir(args.workdir)
    import train  # the reference's train.py
    import models.deep_sets as ds
    assert "pcc_b200" in ds.DeepSets.__module__, "models.deep_sets did not resolve to the B200 package"
    cfg = train.load_config("configs/base.yaml", f"configs/{args.model}.yaml")
    cfg["trainer"]["epochs"] = args.epochs
    import torch
    log_dir = train.train_model(args.model, args.dataset, cfg, return_log_dir=True)
    sd = torch.load(os.path.join(log_dir, "model.pt"), map_location="cpu")
    print(f"REFERENCE_TRAIN_OK model={args.model} dataset={args.dataset} device={'cuda' if torch.cuda.is_available() else 'cpu'} "
          f"log_dir={log_dir} state_dict_keys={len(sd)} first_key={next(iter(sd))}")


if __name__ == "__main__":
    main()
