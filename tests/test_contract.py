"""CPU: drop-in contract — constructor kwargs, state_dict layout, error behaviour, and the
C-ABI library exports every symbol include/pcc.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from helpers import golden_cases, load_golden

import pcc_b200
from pcc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pcc.h")).read()
    declared = set(re.findall(r"\b(pcc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"libpcc.so does not export {name}"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.pcc_version() >= 100


@pytest.mark.parametrize("name", golden_cases())
def test_state_dict_layout_equals_reference(name):
    g = load_golden(name)
    m = pcc_b200.DeepSets(**g["cfg"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["sd"].keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(g["sd"][k].shape), k
    m.load_state_dict(g["sd"])  # strict round trip
    for k, v in m.state_dict().items():
        assert torch.equal(v, g["sd"][k])


def test_ctor_signature_and_errors():
    with pytest.raises(ValueError, match="pooling must be"):
        pcc_b200.DeepSets(3, [8], [8], 1, "relu", pooling="median")
    m = pcc_b200.DeepSets(3, [8], [8], 1, "relu", sparse_batching=False)  # accepted, ignored
    assert m.sparse_batching is False and m.phi_output_dim == 8
    with pytest.raises(AttributeError):  # unknown activation: reference leaves the attribute unset
        pcc_b200.DeepSets(3, [8], [8], 1, "swish")
    with pytest.raises(TypeError):
        pcc_b200.DeepSets(3, [8], [8], 1, "relu", not_a_kwarg=1)


def test_no_cpu_fallback():
    m = pcc_b200.DeepSets(3, [8], [8], 1, "relu", layer_norm=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 3), torch.zeros(4, dtype=torch.long))
    g = pcc_b200.GraphNet(4, 16, 1, "tanh")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.randn(4, 4), torch.zeros(4, dtype=torch.long), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pcc_b200.FusedAdam(m.parameters(), lr=1e-3)


def test_graphnet_state_dict_and_out_of_scope_branches():
    g = pcc_b200.GraphNet(input_dim=4, output_dim=1, hidden_dim=128, activation="tanh", use_gat=False, gat_heads=4,
                          sag_pool=False, pool_ratio=0.5, local_pooling="add", global_pooling="mean",
                          deepchem_style=True)
    keys = set(g.state_dict().keys())
    for conv in ("conv1", "conv2"):
        assert {f"{conv}.lin_rel.weight", f"{conv}.lin_rel.bias", f"{conv}.lin_root.weight"} <= keys
        assert f"{conv}.lin_root.bias" not in keys
    for bn in ("bn1", "bn2", "bn3"):
        assert {f"{bn}.weight", f"{bn}.bias", f"{bn}.running_mean", f"{bn}.running_var",
                f"{bn}.num_batches_tracked"} <= keys
    assert sum(p.numel() for p in g.parameters()) == 68353
    assert tuple(g.fc1.weight.shape) == (256, 128)
    with pytest.raises(NotImplementedError):
        pcc_b200.GraphNet(4, 128, 1, "tanh", use_gat=True)
    with pytest.raises(NotImplementedError):
        pcc_b200.GraphNet(4, 128, 1, "tanh", sag_pool=True)


def test_yaml_default_param_count():
    m = pcc_b200.DeepSets(input_dim=6, phi_layers=[256, 256], rho_layers=[256], output_dim=1, sparse_batching=True,
                          pooling="mean", layer_norm=False, activation="gelu", residual_block=True)
    assert sum(p.numel() for p in m.parameters()) == 199425
    assert list(m.state_dict().keys()) == ["phi.0.weight", "phi.0.bias", "phi.2.linear.weight", "phi.2.linear.bias",
                                           "phi.3.weight", "phi.3.bias", "rho.0.weight", "rho.0.bias",
                                           "rho.2.weight", "rho.2.bias"]


def test_recorded_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/r2 (what bench.py printed on a B200 at the end of the round) carry every
    key of the bench contract: base line, e2e with its copy sizes, roofline, cpu_baseline, launches, clocks; the reference
    arm marks itself and reports zero copies."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "r2", "bench_r2*.json")))
    assert files, "no recorded headline bench line"
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["roofline"]["bound"] in ("hbm", "tensor")
    for k in ("achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-6
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    refs = sorted(glob.glob(os.path.join(root, "profiles", "r2", "bench_ref_r2*.json")))
    r = json.loads(open(refs[-1]).read().strip().splitlines()[-1])
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["config"]["workload"] == d["config"]["workload"]
