"""CPU: the import seam of the reference (train.py:8-9).  With this repo's package directory ahead of
the reference checkout on sys.path, `models.deep_sets` / `models.graph_net` resolve to the B200 modules
while `models.wrapper`, `train`, `utils.*` resolve to the unmodified reference, and train.get_model builds
our DeepSets from the reference's own YAML.  Skipped where the reference checkout does not exist (GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PCC_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference checkout not present")
def test_reference_train_resolves_to_b200_modules(tmp_path):
    code = textwrap.dedent(f"""
        import os, sys
        sys.dont_write_bytecode = True
        sys.path[:0] = [{os.path.join(ROOT, 'point-cloud-classifier_b200')!r}, {REF!r}]
        sys.path.insert(0, {os.path.join(ROOT, 'tools')!r})
        import run_reference_train as R
        R.stub_missing(["h5py", "matplotlib", "seaborn", "torch_geometric"])
        os.chdir({REF!r})
        import train                                    # the reference's train.py, unmodified
        import models.deep_sets as ds, models.graph_net as gn, models.wrapper as mw
        assert "pcc_b200" in ds.DeepSets.__module__ and "pcc_b200" in gn.GraphNet.__module__
        assert mw.__file__.startswith({REF!r}) and train.__file__.startswith({REF!r})
        cfg = train.load_config("configs/base.yaml", "configs/deep_sets.yaml")
        cfg["logging"]["log_dir"] = {str(tmp_path)!r}
        model = train.get_model("deep_sets", cfg)   # ModelWrapper(DeepSets(**cfg["model"]), ...)
        assert type(model.model).__module__.startswith("pcc_b200")
        assert list(model.model.state_dict().keys())[:2] == ["phi.0.weight", "phi.0.bias"]
        gcfg = train.load_config("configs/base.yaml", "configs/graph_net.yaml")
        gcfg["logging"]["log_dir"] = {str(tmp_path)!r}
        gmodel = train.get_model("graph_net", gcfg)
        assert type(gmodel.model).__module__.startswith("pcc_b200")
        print("OK")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
