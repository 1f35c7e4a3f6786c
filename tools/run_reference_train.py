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
    os.chdir(args.workdir)
    import train  # the reference's train.py
    import models.deep_sets as ds
    assert "pcc_b200" in ds.DeepSets.__module__, "models.deep_sets did not resolve to the B200 package"
    cfg = train.load_config("configs/base.yaml", f"configs/{args.model}.yaml")
    cfg["trainer"]["epochs"] = args.epochs
    train.train_model(args.model, args.dataset, cfg)


if __name__ == "__main__":
    main()
