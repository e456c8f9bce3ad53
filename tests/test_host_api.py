"""CPU-side checks of the reference-facing host layer: data generation parity, config schema, module surface,
C-ABI export list, and 'no CPU fallback' behaviour."""
import copy
import ctypes
import hashlib
import json
import os
import re
from collections import defaultdict

import numpy as np
import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(kind, name):
    with open(os.path.join(ROOT, "config_files", kind, f"{name}.yml")) as f:
        return yaml.safe_load(f)


with open(os.path.join(ROOT, "tests", "golden", "scenario_hashes.json")) as _f:
    HASHES = json.load(_f)


@pytest.mark.parametrize("setting", sorted(HASHES))
def test_scenario_data_bit_identical_to_reference(setting):
    """Same YAML + seeds => every tensor of Scenario.get_data() has the sha256 the unmodified reference produced
    (tests/golden/export_configs.py), including the in-place demand-seed shift of the one-store settings."""
    from neural_inventory_control_b200.data_handling import Scenario
    s = copy.deepcopy(_cfg("settings", setting))
    obs = defaultdict(lambda: None, s["observation_params"])
    sc = Scenario(60, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], 64, obs,
                  s["seeds"])
    data = sc.get_data()
    want = HASHES[setting]
    assert sorted(data) == sorted(k for k in want if not k.startswith("__"))
    for k, v in data.items():
        assert v.dtype == torch.float32
        assert list(v.shape) == want[k][0], k
        assert hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest()[:16] == want[k][1], k
    assert s["seeds"] == want["__seeds_after__"]


def test_dataset_split_dev_first():
    from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
    s = copy.deepcopy(_cfg("settings", "one_store_lost"))
    obs = defaultdict(lambda: None, s["observation_params"])
    sc = Scenario(20, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], 48, obs,
                  s["seeds"])
    train, dev = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True, sample_index_for_split=16)
    full = sc.get_data()
    assert len(dev) == 16 and len(train) == 32
    assert torch.equal(dev.data["demands"], full["demands"][:16])
    assert torch.equal(train.data["demands"], full["demands"][16:])
    item = train[3]
    assert set(item) == set(full) and item["demands"].shape == (1, 20)


@pytest.mark.parametrize("setting,policy,keys", [
    ("one_store_lost", "vanilla_one_store", ["net.master.0.weight", "net.master.6.bias"]),
    ("serial_system", "vanilla_serial", ["net.master.0.weight", "net.master.4.bias"]),
    ("one_warehouse_lost_demand", "vanilla_warehouse", ["net.master.0.weight", "net.master.6.bias"]),
    ("one_warehouse_lost_demand", "symmetry_aware", ["net.context.0.weight", "net.store.4.bias", "net.warehouse.4.bias"]),
])
def test_policy_construction_and_state_dict_names(setting, policy, keys):
    """Same ModuleDict / Sequential layout as the reference => same state_dict keys; Lazy layers materialise on CPU
    with plain torch (policy forward is torch code; only Simulator / fused rollouts need the GPU)."""
    from neural_inventory_control_b200.data_handling import Scenario
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    s = copy.deepcopy(_cfg("settings", setting))
    p = _cfg("policies_and_hyperparams", policy)
    obs = defaultdict(lambda: None, s["observation_params"])
    sc = Scenario(10, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], 8, obs,
                  s["seeds"])
    torch.manual_seed(0)
    model = NeuralNetworkCreator().create_neural_network(sc, p["nn_params"], device="cpu")
    data = sc.get_data()
    observation = {"store_inventories": data["initial_inventories"]}
    if s["problem_params"]["n_warehouses"] > 0:
        observation["warehouse_inventories"] = data["initial_warehouse_inventories"]
    if s["problem_params"]["n_extra_echelons"] > 0:
        observation["echelon_inventories"] = data["initial_echelon_inventories"]
    for k in ("mean", "std", "underage_costs", "lead_times"):
        if k in data:
            observation[k] = data[k]
    assert model.fusable_spec() is None  # still lazy
    action = model(observation)
    S, W = s["problem_params"]["n_stores"], max(s["problem_params"]["n_warehouses"], 1)
    assert action["stores"].shape == (8, S, W)
    sd = model.state_dict()
    for k in keys:
        assert k in sd
    spec = model.fusable_spec()
    assert spec is not None and spec.arch == p["nn_params"]["name"]
    if "warehouse_upper_bound_mult" in p["nn_params"]:
        mean = sc.store_params["demand"]["mean"]
        total = float(np.sum(mean)) if not isinstance(mean, float) else mean
        assert abs(spec.warehouse_upper_bound - p["nn_params"]["warehouse_upper_bound_mult"] * total) < 1e-3


def test_root_modules_star_export_surface():
    import subprocess
    import sys
    code = ("from trainer import *\n"
            "names = ['torch','nn','np','pd','DataLoader','Dataset','DefaultDict','copy','datetime','os','Scenario',"
            "'DatasetCreator','MyDataset','NeuralNetworkCreator','PolicyLoss','Simulator','Trainer','VanillaOneStore',"
            "'VanillaSerial','VanillaWarehouse']\n"
            "missing = [n for n in names if n not in globals()]\n"
            "assert not missing, missing\nprint('ok')")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_c_abi_exports_every_declared_symbol():
    """libhdpo_b200.so (built by __graft_entry__.build()) loads and exports every function include/hdpo_b200.h
    declares; no compute call is made here (no GPU)."""
    from neural_inventory_control_b200 import _capi, _lib
    header = open(os.path.join(ROOT, "include", "hdpo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hdpo_[a-z_0-9]+)\s*\(", header)))
    assert set(declared) == set(_capi.EXPORTS), (declared, _capi.EXPORTS)
    if not os.path.exists(_lib.lib_path()):
        pytest.skip("libhdpo_b200.so not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hdpo_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    from neural_inventory_control_b200 import _capi as K
    assert ctypes.sizeof(K.Problem) == 10 * 4
    assert ctypes.sizeof(K.Statics) == 10 * 8
    assert ctypes.sizeof(K.Mlp) == 4 * (1 + 9 + 2)
    # HdpoRolloutDesc: Problem (40) + 10 int32 + 2 float + 3 Mlp (48 each) + pointer (8-aligned)
    # + ABI 2: 2 int32 + float + int32 + 2 uint64 + 2 pointers (Philox demand source)
    assert ctypes.sizeof(K.RolloutDesc) == 40 + 40 + 8 + 3 * 48 + 8 + 16 + 16 + 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback():
    from neural_inventory_control_b200 import engine
    from neural_inventory_control_b200.environment import Simulator
    s = copy.deepcopy(_cfg("settings", "one_store_lost"))
    obs = defaultdict(lambda: None, s["observation_params"])
    from neural_inventory_control_b200.data_handling import Scenario
    sc = Scenario(10, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], 4, obs,
                  s["seeds"])
    with pytest.raises(RuntimeError):
        Simulator(device="cpu").reset(5, s["problem_params"], sc.get_data(), obs)
    with pytest.raises(RuntimeError):
        engine.FusedRollout(engine.PolicySpec("vanilla_one_store", ([4, 32, 1], "elu", None)), s["problem_params"],
                            sc.get_data(), 5)
