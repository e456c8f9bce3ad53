"""The reference's shipped weight-shared policy (GNN, /root/reference/neural_networks.py:742-1492) on this engine.

It runs on the generic per-step path (torch policy + one K3 simulator kernel per period). Pins: fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py, kind "gnn"): the first action (policy alone, CPU), and the 50-period
costs + parameter gradients through Trainer.simulate_batch on the GPU."""
import copy
import os
from collections import defaultdict

import numpy as np
import pytest
import torch
import yaml

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["one_warehouse", "many_warehouses", "transshipment"]  # the last one: gnn_transshipment.yml (no hold option)


def _cfg(kind, name):
    with open(os.path.join(ROOT, "config_files", kind, f"{name}.yml")) as f:
        return yaml.safe_load(f)


class _Scenario:
    def __init__(self, problem_params):
        self.problem_params, self.store_params = problem_params, {"demand": {"mean": [1.0]}}


def _model(meta, g, device, dtype=torch.float32):
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    nn_params = copy.deepcopy(_cfg("policies_and_hyperparams", meta["policy"]))["nn_params"]
    model = NeuralNetworkCreator().create_neural_network(_Scenario(meta["problem_params"]), nn_params, device=device)
    return model


def _observation(data, device, dtype):
    """What Simulator.reset hands the policy in period 0 (environment.py:301-345), built by hand for the CPU test."""
    t = lambda k: torch.tensor(data[k], device=device, dtype=dtype)  # noqa: E731
    obs = {"store_inventories": t("initial_inventories"), "warehouse_inventories": t("initial_warehouse_inventories"),
           "warehouse_lead_times": t("warehouse_lead_times"), "warehouse_holding_costs": t("warehouse_holding_costs"),
           "holding_costs": t("holding_costs"), "underage_costs": t("underage_costs"), "lead_times": t("lead_times"),
           "mean": t("mean"), "std": t("std")}
    if "warehouse_edge_costs" in data:
        obs["warehouse_edge_costs"] = t("warehouse_edge_costs")
    return obs


@pytest.mark.parametrize("name", CASES)
def test_gnn_first_action_matches_reference(name):
    meta, g = G.load("gnn", name)
    assert meta["nn_name"] == "gnn"
    for dtype, tag, tol in ((torch.float32, "ref", 2e-5), (torch.float64, "ref", 2e-6)):
        model = _model(meta, g, "cpu")
        obs = _observation(g["data"], "cpu", torch.float32)
        with torch.no_grad():
            model(obs)  # materialise the lazy layers
        model.load_state_dict({k: torch.tensor(v) for k, v in g["param"].items()})
        # state_dict names are the reference's
        assert set(model.state_dict()) == set(g["param"])
        if dtype == torch.float64:
            model = model.double()
            obs = _observation(g["data"], "cpu", torch.float64)
        with torch.no_grad():
            act = model(obs)
        for k, want in g[tag].items():
            if not k.startswith("action0/"):
                continue
            got = act[k.split("/", 1)[1]].numpy()
            assert got.shape == want.shape, (k, got.shape, want.shape)
            assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), (k, np.abs(got - want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gnn_rollout_costs_and_gradients_match_reference(name):
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.trainer import Trainer
    dev = "cuda:0"
    meta, g = G.load("gnn", name)
    model = _model(meta, g, dev)
    data = {k: torch.tensor(v, device=dev) for k, v in g["data"].items()}
    pp = meta["problem_params"]
    obs_params = defaultdict(lambda: None, copy.deepcopy(_cfg("settings", meta["setting"])["observation_params"]))
    if "mean" in g["data"]:  # the transshipment fixture was generated with the demand moments switched on (the GNN reads them)
        obs_params["include_static_features"].update(mean=True, std=True)
    tr, sim = Trainer(device=dev), Simulator(device=dev)
    with torch.no_grad():
        tr.simulate_batch(PolicyLoss(), sim, model, 1, pp, {k: v[:2] for k, v in data.items()}, obs_params)
    assert tr.last_path == "generic"
    model.load_state_dict({k: torch.tensor(v) for k, v in g["param"].items()})

    class Capture(PolicyLoss):  # same interface the reference hands the reward to (trainer.py:206)
        def __init__(self):
            super().__init__()
            self.rewards = []

        def forward(self, observation, action, reward):
            self.rewards.append(reward.detach().clone())
            return reward.sum()
    loss = Capture()
    T, ignore = meta["T"], meta["ignore_periods"]
    total, report = tr.simulate_batch(loss, sim, model, T, pp, data, obs_params, ignore)
    B = data["demands"].shape[0]
    (total / (B * T * pp["n_stores"])).backward()
    ref, ref64 = g["ref"], g["ref64"]
    true_tb = ref64["reward_tb"]
    mine_tb = torch.stack(loss.rewards, 0).double().cpu().numpy()
    true_b = true_tb.sum(0)
    floor = np.abs(ref["reward_tb"].astype(np.float64).sum(0) / true_b - 1).max()
    tol = max(1e-5, 3 * floor)
    assert np.abs(mine_tb.sum(0) / true_b - 1).max() <= tol, (np.abs(mine_tb.sum(0) / true_b - 1).max(), floor)
    assert abs(float(total) - float(ref64["total"])) <= tol * abs(float(ref64["total"]))
    assert abs(float(report) - float(ref64["report"])) <= tol * abs(float(ref64["total"]))
    keys = sorted(k for k, _ in model.named_parameters())
    mine = np.concatenate([dict(model.named_parameters())[k].grad.detach().cpu().numpy().ravel() for k in keys])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in keys])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in keys])
    gfloor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, 3 * gfloor), (G.rel_l2(mine, r64), gfloor)


@pytest.mark.gpu
def test_main_run_trains_gnn(monkeypatch, capsys):
    """`python main_run.py train one_warehouse_lost_demand gnn` - the command that was a KeyError in round 1."""
    import main_run
    real_load = main_run.load_yaml

    def small(path):
        cfg = real_load(path)
        if "params_by_dataset" in cfg:
            cfg["params_by_dataset"]["train"].update(n_samples=256, batch_size=128)
            cfg["params_by_dataset"]["dev"].update(n_samples=128, batch_size=128)
        if "trainer_params" in cfg:
            cfg["trainer_params"].update(epochs=2, do_dev_every_n_epochs=1, save_model=False)
        return cfg
    monkeypatch.setattr(main_run, "load_yaml", small)
    monkeypatch.chdir(ROOT)
    main_run.main(["main_run.py", "train", "one_warehouse_lost_demand", "gnn"])
    out = capsys.readouterr().out
    assert "Average per-period train loss" in out
