"""Real-data observation features and the policies that consume them.

Reference: environment.py:436-501 (past-demand window, time features, period shift, profit objective),
neural_networks.py:430-740 (DataDrivenNet, quantile policies, JustInTime), quantile_forecaster.py. The policies run on
the generic per-step path (torch policy + one K3 simulator kernel per period). Pins: fixtures produced by the UNMODIFIED
reference on the shipped Favorita files (tests/golden/make_golden.py, kind "realdata"): the first action (policy +
observation glue, CPU) and the 50-period costs + parameter gradients through Trainer.simulate_batch on the GPU."""
import copy
import os
from collections import defaultdict

import numpy as np
import pytest
import torch
import yaml

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAINABLE = ["many_warehouses_data_driven", "one_store_transformed_nv", "one_store_fixed_quantile"]
BENCHMARKS = ["many_warehouses_just_in_time", "one_store_quantile_nv", "one_store_returns_nv", "one_store_just_in_time"]


def _cfg(kind, name):
    with open(os.path.join(ROOT, "config_files", kind, f"{name}.yml")) as f:
        return yaml.safe_load(f)


class _Scenario:
    def __init__(self, problem_params):
        self.problem_params, self.store_params = problem_params, {"demand": {"mean": [1.0]}}


def _model(meta, g, device, tmp_path):
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    nn_params = copy.deepcopy(_cfg("policies_and_hyperparams", meta["policy"]))["nn_params"]
    if "aux" in g:  # the frozen forecaster's weights travel inside the fixture (the reference ships them as a .pt file)
        path = os.path.join(str(tmp_path), "forecaster.pt")
        torch.save({k.split("/", 1)[1]: torch.tensor(v) for k, v in g["aux"].items()}, path)
        nn_params["forecaster_location"] = path
    return NeuralNetworkCreator().create_neural_network(_Scenario(meta["problem_params"]), nn_params, device=device)


def _load_params(model, g):
    if "param" in g:
        model.load_state_dict({k: torch.tensor(v) for k, v in g["param"].items()}, strict=False)


def _first_observation(meta, data, obs_params, device, dtype):
    """Period-0 observation through the package's own feature glue (Simulator.initialize_observation), without the
    CUDA-only parts of reset()."""
    from neural_inventory_control_b200.environment import Simulator
    sim = Simulator(device=device)
    d = {k: torch.tensor(v, device=device, dtype=dtype) for k, v in data.items()}
    sim.problem_params, sim.observation_params = meta["problem_params"], obs_params
    sim.batch_size, sim.n_stores = len(d["initial_inventories"]), meta["problem_params"]["n_stores"]
    sim._internal_data = {"demands": d["demands"], "period_shift": obs_params["demand"]["period_shift"]}
    for kind in ("time_features", "sample_features"):
        if obs_params[kind] is not None:
            sim._internal_data.update({k: d[k] for k in obs_params[kind]})
    obs = sim.initialize_observation(d, obs_params)
    obs["internal_data"] = sim._internal_data
    return obs


@pytest.mark.parametrize("name", TRAINABLE + BENCHMARKS)
def test_first_action_matches_reference(name, tmp_path):
    meta, g = G.load("realdata", name)
    obs_params = defaultdict(lambda: None, _cfg("settings", meta["setting"])["observation_params"])
    assert obs_params["demand"]["past_periods"] == 16 and obs_params["demand"]["period_shift"] == 16
    model = _model(meta, g, "cpu", tmp_path)
    obs = _first_observation(meta, g["data"], obs_params, "cpu", torch.float32)
    # the window the policy sees in period 0 is the 16 periods before the shifted start
    assert torch.equal(obs["past_demands"], torch.tensor(g["data"]["demands"][:, :, :16]))
    assert torch.equal(obs["days_from_christmas"], torch.tensor(g["data"]["days_from_christmas"][:, :, 16]))
    with torch.no_grad():
        model(obs)  # materialise the lazy layers
    _load_params(model, g)
    if "param" in g:
        mine = {k for k, v in model.state_dict().items() if not isinstance(v, torch.nn.parameter.UninitializedTensorMixin)}
        assert mine == set(g["param"])  # state_dict names are the reference's (the forecaster is not part of it)
    with torch.no_grad():
        act = model(obs)
    n = 0
    for k, want in g["ref"].items():
        if not k.startswith("action0/"):
            continue
        got = act[k.split("/", 1)[1]].numpy()
        assert got.shape == want.shape, (k, got.shape, want.shape)
        assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), (k, np.abs(got - want).max())
        n += 1
    assert n >= 1


def test_past_demand_window_is_left_padded_with_zeros():
    """environment.py:436-458: before `past_periods` periods have elapsed the window is zero-filled on the left."""
    from neural_inventory_control_b200.environment import Simulator
    sim = Simulator(device="cpu")
    sim._internal_data = {"period_shift": 0}
    data = {"demands": torch.arange(2 * 3 * 10, dtype=torch.float32).reshape(2, 3, 10)}
    op = {"demand": {"past_periods": 4, "period_shift": 0}}
    assert torch.equal(sim.update_past_demands(data, op, 2, 3, 0), torch.zeros(2, 3, 4))
    w = sim.update_past_demands(data, op, 2, 3, 2)
    assert torch.equal(w[:, :, :2], torch.zeros(2, 3, 2)) and torch.equal(w[:, :, 2:], data["demands"][:, :, :2])
    assert torch.equal(sim.update_past_demands(data, op, 2, 3, 7), data["demands"][:, :, 3:7])


def test_forecaster_interpolates_between_predicted_quantiles():
    """quantile_forecaster.py:61-101: get_quantile is the piecewise-linear inverse CDF through the predicted quantiles,
    extended linearly to probabilities 0 and 1; get_implied_percentile inverts it."""
    from neural_inventory_control_b200.quantile_forecaster import FullyConnectedForecaster
    torch.manual_seed(3)
    f = FullyConnectedForecaster([8], lead_times=[4, 5, 6])
    x = torch.rand(5, 2, 17)
    lt = torch.randint(4, 7, (5, 2)).float()
    with torch.no_grad():
        # make the predicted quantiles strictly increasing so that the inverse is unique
        f(x)
        f.net[-1].weight.zero_()
        f.net[-1].bias.copy_(torch.arange(1, 58, dtype=torch.float32).reshape(19, 3).flatten())
        grid = f(x)  # [5,2,19,3], constant rows: quantile i of lead-time slot j = 3 i + j + 1
        assert grid.shape == (5, 2, 19, 3)
        q = torch.full((5, 2), 0.5)
        level = f.get_quantile(x, q, lt)
        want = 3 * 9 + (lt - 4) + 1  # q = 0.5 is grid point 9
        assert torch.allclose(level, want, atol=1e-4)
        q2 = torch.full((5, 2), 0.525)  # half-way between grid points 9 and 10
        assert torch.allclose(f.get_quantile(x, q2, lt), want + 1.5, atol=1e-3)
        back = f.get_implied_percentile(x, lt, want + 1.5)
        assert torch.allclose(back, q2, atol=1e-4)
        # below the first predicted quantile: linear extrapolation towards probability 0
        assert torch.allclose(f.get_quantile(x, torch.full((5, 2), 0.025), lt), (lt - 4) + 1 - 1.5, atol=1e-3)


def _rollout(name, tmp_path, need_grad):
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.trainer import Trainer
    dev = "cuda:0"
    meta, g = G.load("realdata", name)
    model = _model(meta, g, dev, tmp_path)
    data = {k: torch.tensor(v, device=dev) for k, v in g["data"].items()}
    pp = meta["problem_params"]
    obs_params = defaultdict(lambda: None, _cfg("settings", meta["setting"])["observation_params"])
    tr, sim = Trainer(device=dev), Simulator(device=dev)
    with torch.no_grad():
        tr.simulate_batch(PolicyLoss(), sim, model, 1, pp, {k: v[:2] for k, v in data.items()}, obs_params)
    assert tr.last_path == "generic"
    _load_params(model, g)

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
    if need_grad:
        (total / (B * T * pp["n_stores"])).backward()
    ref, ref64 = g["ref"], g["ref64"]
    true_tb = ref64["reward_tb"]
    mine_tb = torch.stack(loss.rewards, 0).double().cpu().numpy()
    scale = np.abs(true_tb).sum(0)  # profit objective: per-scenario totals can be near zero, compare against the turnover
    floor = (np.abs(ref["reward_tb"].astype(np.float64).sum(0) - true_tb.sum(0)) / scale).max()
    tol = max(1e-5, 3 * floor)
    err = (np.abs(mine_tb.sum(0) - true_tb.sum(0)) / scale).max()
    assert err <= tol, (err, floor)
    assert abs(float(total) - float(ref64["total"])) <= tol * np.abs(true_tb).sum()
    assert abs(float(report) - float(ref64["report"])) <= tol * np.abs(true_tb).sum()
    return model, ref, ref64


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRAINABLE)
def test_rollout_costs_and_gradients_match_reference(name, tmp_path):
    model, ref, ref64 = _rollout(name, tmp_path, need_grad=True)
    keys = sorted(k for k, p in model.named_parameters() if p.grad is not None)
    assert keys == sorted(k.split("/", 1)[1] for k in ref if k.startswith("grad/"))
    mine = np.concatenate([dict(model.named_parameters())[k].grad.detach().cpu().numpy().ravel() for k in keys])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in keys])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in keys])
    gfloor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, 3 * gfloor), (G.rel_l2(mine, r64), gfloor)


@pytest.mark.gpu
@pytest.mark.parametrize("name", BENCHMARKS)
def test_benchmark_policy_rollout_costs_match_reference(name, tmp_path):
    _rollout(name, tmp_path, need_grad=False)


# ---------------------------------------------------------------------------------------------------------------------
# dataset pipeline on real-data FILES (data_handling.py:9-123, 162-168, 398-458): fixture files + hashes of what the
# unmodified reference makes of them (tests/golden/make_realdata_pipeline_golden.py)
# ---------------------------------------------------------------------------------------------------------------------
def _pipeline_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_realdata_pipeline_golden", os.path.join(ROOT, "tests", "golden", "make_realdata_pipeline_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_real_data_pipeline_matches_reference():
    """Scenario reads the demand file + calendar file, DatasetCreator splits BY PERIOD: every tensor of the full data
    and of the train / dev / test windows has the sha256 the unmodified reference produced from the same files."""
    import json
    from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
    mod = _pipeline_module()
    with open(os.path.join(ROOT, "tests", "golden", "realdata_pipeline_hashes.json")) as f:
        want = json.load(f)
    got = mod.run(Scenario, DatasetCreator)
    for part in ("full", "train", "dev", "test"):
        assert sorted(got[part]) == sorted(want[part]), part
        for k, v in want[part].items():
            assert got[part][k] == v, (part, k)
    # the windows are slices of the full series, the per-sample tensors are shared
    assert want["train"]["demands"][0] == [20, 3, 40] and want["dev"]["demands"][0] == [20, 3, 24]


@pytest.mark.gpu
def test_main_run_trains_data_driven_net_on_real_data_files(monkeypatch, capsys):
    """`python main_run.py train <real-data setting> data_driven_net`: split by period, past-demand window + days from
    Christmas in the observation, profit objective, DataDrivenNet on the generic per-step path - two epochs on the
    fixture files."""
    import copy as _copy
    import main_run
    mod = _pipeline_module()
    real_load = main_run.load_yaml

    def load(path):
        if "settings" in path:
            s = _copy.deepcopy(mod.setting_with_paths())
            s["test_seeds"] = _copy.deepcopy(s["seeds"])
            s["params_by_dataset"] = {
                "train": {"n_samples": 24, "batch_size": 12, "periods": 32, "ignore_periods": 8},
                "dev": {"n_samples": 24, "batch_size": 24, "periods": 16, "ignore_periods": 8},
                "test": {"n_samples": 24, "batch_size": 24, "periods": 16, "ignore_periods": 8}}
            return s
        cfg = real_load(path)
        cfg["trainer_params"].update(epochs=2, do_dev_every_n_epochs=1, save_model=False)
        return cfg
    monkeypatch.setattr(main_run, "load_yaml", load)
    monkeypatch.chdir(ROOT)
    main_run.main(["main_run.py", "train", "fixture_real_data", "data_driven_net"])
    out = capsys.readouterr().out
    assert "Average per-period train loss" in out
