"""GPU tests of the reference-facing Python layer (Trainer.simulate_batch / Simulator.step / policies) against the
reference-generated goldens: the same state_dict and the same Scenario tensors go through
  (a) the FUSED path  (one forward kernel + one adjoint kernel per batch), and
  (b) the GENERIC path (torch policy + one K3 kernel per period, torch autograd across periods),
and both must reproduce the reference's totals and parameter gradients."""
import copy
import os
from collections import defaultdict

import numpy as np
import pytest
import torch
import yaml

import golden_util as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(kind, name):
    with open(os.path.join(ROOT, "config_files", kind, f"{name}.yml")) as f:
        return yaml.safe_load(f)


class _FakeScenario:
    def __init__(self, problem_params, store_params):
        self.problem_params, self.store_params = problem_params, store_params


def _build(meta, g, device):
    """Model from this repo's NeuralNetworkCreator carrying the golden weights, plus the golden batch on device."""
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    p = copy.deepcopy(_cfg("policies_and_hyperparams", meta["policy"]))
    nn_params = p["nn_params"]
    nn_params["neurons_per_hidden_layer"] = meta["neurons_per_hidden_layer"]
    pp = meta["problem_params"]
    scen = _FakeScenario(pp, {"demand": {"mean": [meta["warehouse_upper_bound"] / max(nn_params.get(
        "warehouse_upper_bound_mult", 1), 1)]}})
    model = NeuralNetworkCreator().create_neural_network(scen, nn_params, device=device)
    data = {k: torch.tensor(v, device=device) for k, v in g["data"].items()}
    return model, data, pp


def _obs_params(meta):
    s = _cfg("settings", meta["setting"])
    return defaultdict(lambda: None, s["observation_params"])


def _run(meta, g, fused, device="cuda:0"):
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.trainer import Trainer
    model, data, pp = _build(meta, g, device)
    tr = Trainer(device=device)
    tr.use_fused = fused
    sim = Simulator(device=device)
    obs_params = _obs_params(meta)
    # materialise lazily-shaped layers, then load the golden weights
    with torch.no_grad():
        tr.simulate_batch(PolicyLoss(), sim, model, 1, pp, {k: v[:2] for k, v in data.items()}, obs_params)
    model.load_state_dict({k: torch.tensor(v) for k, v in g["param"].items()})
    total, report = tr.simulate_batch(PolicyLoss(), sim, model, meta["T"], pp, data, obs_params,
                                      meta["ignore_periods"])
    B = data["demands"].shape[0]
    (total / (B * meta["T"] * pp["n_stores"])).backward()
    grads = {k: v.grad.detach().cpu().numpy() for k, v in model.named_parameters()}
    return tr.last_path, float(total), float(report), grads, sim


CASES = ["one_store_lost", "one_store_backlogged_lead20", "serial_system", "one_warehouse_s5", "many_warehouses_2x10"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "generic"])
def test_simulate_batch_matches_reference(name, fused):
    meta, g = G.load("rollout", name)
    path, total, report, grads, sim = _run(meta, g, fused)
    assert path == ("fused" if fused else "generic")
    ref, ref64 = g["ref"], g["ref64"]
    floor_c = abs(float(ref["total"]) / float(ref64["total"]) - 1)
    tol = max(1e-5, 3 * floor_c)
    assert abs(total / float(ref64["total"]) - 1) <= tol
    assert abs(report / float(ref64["report"]) - 1) <= tol
    keys = sorted(grads)
    mine = np.concatenate([grads[k].ravel() for k in keys])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in keys])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in keys])
    floor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, 3 * floor), (G.rel_l2(mine, r64), floor)
    assert int(sim.observation["current_period"]) == meta["T"]


def test_generic_step_observation_contract():
    """Simulator.reset/step keep the reference's observation dict, int64 shift tables and period bookkeeping."""
    from neural_inventory_control_b200.environment import Simulator
    meta, g = G.load("step", "many_warehouses")
    dev = "cuda:0"
    data = {k: torch.tensor(v, device=dev) for k, v in g["data"].items()}
    obs_params = defaultdict(lambda: None, {
        "include_warehouse_inventory": True,
        "include_static_features": {"holding_costs": True, "underage_costs": True, "lead_times": True},
        "demand": {"past_periods": 0, "period_shift": 1}})
    sim = Simulator(device=dev)
    obs, info = sim.reset(2, meta["problem_params"], data, obs_params)
    assert info is None
    for k in ("store_inventories", "warehouse_inventories", "warehouse_lead_times", "warehouse_holding_costs",
              "warehouse_edge_costs", "holding_costs", "underage_costs", "lead_times", "current_period"):
        assert k in obs, k
    assert obs["current_period"].device.type == "cpu"
    shift = sim._internal_data["allocation_shift"]
    assert shift.dtype == torch.int64 and np.array_equal(shift.cpu().numpy(), g["ref"]["allocation_shift"])
    assert np.array_equal(sim._internal_data["warehouse_allocation_shift"].cpu().numpy(),
                          g["ref"]["warehouse_allocation_shift"])
    action = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in g["action"].items()}
    obs, reward, terminated, _, _ = sim.step(action)
    np.testing.assert_allclose(reward.detach().cpu().numpy(), g["ref"]["reward"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(obs["store_inventories"].detach().cpu().numpy(), g["ref"]["new/store_inventories"],
                               rtol=2e-6, atol=2e-6)
    assert not bool(terminated)
    up = g["up"]
    loss = (reward * torch.tensor(up["reward"], device=dev)).sum() \
        + (obs["store_inventories"] * torch.tensor(up["store_inventories"], device=dev)).sum() \
        + (obs["warehouse_inventories"] * torch.tensor(up["warehouse_inventories"], device=dev)).sum()
    loss.backward()
    for k, v in action.items():
        np.testing.assert_allclose(v.grad.cpu().numpy(), g["ref"][f"grad/action_{k}"], rtol=1e-5, atol=1e-5)
    obs, reward, terminated, _, _ = sim.step({k: v.detach() for k, v in action.items()})
    assert bool(terminated) and int(obs["current_period"]) == 2


def test_training_loop_reduces_loss_and_eval_modes_run():
    """A few epochs of `Trainer.train` through the fused path on the shipped one-store setting (reduced sample
    counts): the reported train loss must go down; dev (no_grad) and test (discrete allocation) epochs must run."""
    from torch.utils.data import DataLoader
    from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    from neural_inventory_control_b200.trainer import Trainer
    dev = "cuda:0"
    s = copy.deepcopy(_cfg("settings", "one_store_lost"))
    p = copy.deepcopy(_cfg("policies_and_hyperparams", "vanilla_one_store"))
    obs_params = defaultdict(lambda: None, s["observation_params"])
    pbd = s["params_by_dataset"]
    pbd["train"].update(n_samples=2048, batch_size=512)
    pbd["dev"].update(n_samples=512, batch_size=512, periods=60, ignore_periods=30)
    pbd["test"].update(n_samples=256, batch_size=256, periods=80, ignore_periods=40)
    common = (s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"])
    sc = Scenario(60, *common, 2048 + 512, obs_params, s["seeds"])
    train, devset = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True, sample_index_for_split=512)
    sc_test = Scenario(80, *common, 256, obs_params, s["test_seeds"])
    test = DatasetCreator().create_datasets(sc_test, split=False)
    loaders = {"train": DataLoader(train, batch_size=512, shuffle=True),
               "dev": DataLoader(devset, batch_size=512, shuffle=False),
               "test": DataLoader(test, batch_size=256, shuffle=False)}
    torch.manual_seed(0)
    model = NeuralNetworkCreator().create_neural_network(sc_test, p["nn_params"], device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=p["optimizer_params"]["learning_rate"])
    tr, sim = Trainer(device=dev), Simulator(device=dev)
    tp = p["trainer_params"]
    tp.update(epochs=12, do_dev_every_n_epochs=4, print_results_every_n_epochs=100, save_model=False)
    tr.train(12, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs_params, pbd, tp)
    assert tr.last_path == "fused"
    assert len(tr.all_train_losses) == 12 and len(tr.all_dev_losses) == 12
    assert tr.all_train_losses[-1] < 0.8 * tr.all_train_losses[0]
    assert tr.best_performance_data["model_params_to_save"] is not None
    loss, report = tr.test(PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs_params, pbd,
                           discrete_allocation=True)
    assert np.isfinite(loss) and np.isfinite(report) and report > 0


@pytest.mark.parametrize("fused", [False, True], ids=["generic", "fused"])
def test_symmetry_aware_trainer_paths_match_oracle(fused):
    """The re-created SymmetryAware policy (absent from the reference snapshot's sources) through
    Trainer.simulate_batch on both paths - GENERIC (torch policy + K3 step kernels + torch autograd) and FUSED (trunk
    GEMMs + rollout_sym.cu heads, one autograd node) - against the float64 oracle restatement of the same recovered
    forward: costs and all three nets' gradients."""
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    from neural_inventory_control_b200.trainer import Trainer
    from oracle import hdpo_oracle as O
    dev = "cuda:0"
    meta, g = G.load("rollout", "one_warehouse_s5")
    p = copy.deepcopy(_cfg("policies_and_hyperparams", "symmetry_aware"))
    p["nn_params"]["neurons_per_hidden_layer"] = {"context": [24], "store": [16, 16], "warehouse": [8, 8]}
    p["nn_params"]["output_sizes"]["context"] = 12
    pp = meta["problem_params"]
    wub = meta["warehouse_upper_bound"]
    scen = _FakeScenario(pp, {"demand": {"mean": [wub / 4]}})
    torch.manual_seed(3)
    model = NeuralNetworkCreator().create_neural_network(scen, p["nn_params"], device=dev)
    data = {k: torch.tensor(v[:16], device=dev) for k, v in g["data"].items()}
    obs_params = _obs_params(meta)
    tr, sim = Trainer(device=dev), Simulator(device=dev)
    tr.use_fused = fused
    T = 10
    total, report = tr.simulate_batch(PolicyLoss(), sim, model, T, pp, data, obs_params, 3)
    assert tr.last_path == ("fused" if fused else "generic")
    B = 16
    (total / (B * T * pp["n_stores"])).backward()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    acts = p["nn_params"]
    nets = {m: O.mlp_from_state_dict(sd, m, acts["inner_layer_activations"][m], acts["output_layer_activation"][m]
                                     ).astype(np.float64) for m in ("context", "store", "warehouse")}
    pol = O.Policy("symmetry_aware", nets, float(model.warehouse_upper_bound[0]), prop_eps=1e-15)
    d64 = {k: v[:16].astype(np.float64) for k, v in g["data"].items()}
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), d64, T)
    assert abs(float(total) / fwd["total"] - 1) < 1e-5
    assert abs(float(report) / fwd["reward_tb"][3:].sum() - 1) < 1e-5
    flat = O.flatten_grads(pol, grads)
    for k, v in model.named_parameters():
        assert G.rel_l2(v.grad.cpu().numpy(), flat[k]) < 5e-5, k


def test_device_resident_batches_match_dataloader_semantics():
    """DeviceBatches: same batch sizes / last partial batch / coverage as the torch DataLoader it stands in for; a
    shuffled epoch is a permutation of the dataset; gathered rows are bit-identical to the source rows."""
    from torch.utils.data import DataLoader
    from neural_inventory_control_b200.data_handling import MyDataset
    from neural_inventory_control_b200.device_dataset import DeviceBatches, eligible
    n = 1000
    g = torch.Generator().manual_seed(0)
    data = {"demands": torch.rand(n, 3, 17, generator=g), "initial_inventories": torch.rand(n, 3, 4, generator=g),
            "holding_costs": torch.rand(n, 3, generator=g), "tag": torch.arange(n).float()}
    ds = MyDataset(n, data)
    for shuffle in (False, True):
        loader = DataLoader(ds, batch_size=256, shuffle=shuffle)
        assert eligible(loader)
        db = DeviceBatches(loader, "cuda:0")
        assert len(db) == len(loader) == 4 and len(db.dataset) == n
        seen = []
        for batch in db:
            assert set(batch) == set(data)
            tags = batch["tag"].long().cpu()
            seen.append(tags)
            for k in data:
                assert torch.equal(batch[k].cpu(), data[k][tags])
        sizes = [len(t) for t in seen]
        assert sizes == [256, 256, 256, 232]
        allt = torch.cat(seen)
        assert torch.equal(allt.sort().values, torch.arange(n))
        if shuffle:
            assert not torch.equal(allt, torch.arange(n))
            again = torch.cat([b["tag"].long().cpu() for b in db])
            assert not torch.equal(again, allt)  # fresh permutation per epoch
        else:
            assert torch.equal(allt, torch.arange(n))
    assert len(DeviceBatches(DataLoader(ds, batch_size=256, shuffle=True, drop_last=True), "cuda:0")) == 3


def test_main_run_cli_train_and_long_horizon_test(monkeypatch, capsys):
    """`python main_run.py train|test one_store_lost vanilla_one_store` exactly as the reference wires it (main_run.py:
    argv, YAML keys, datasets incl. the 32768 x 5000-period test set). `train` is capped at 2 epochs here; `test` runs
    the long-horizon evaluation (trainer.py:121-141: 5000 periods, 3000 ignored, discrete allocation for Poisson demand)
    on the fused forward-only rollout."""
    import main_run
    from neural_inventory_control_b200.trainer import Trainer
    monkeypatch.chdir(ROOT)
    paths = []
    orig_train, orig_sim = Trainer.train, Trainer.simulate_batch

    def capped_train(self, epochs, *a, **k):
        return orig_train(self, min(epochs, 2), *a, **k)

    def spy(self, *a, **k):
        out = orig_sim(self, *a, **k)
        paths.append(self.last_path)
        return out

    monkeypatch.setattr(Trainer, "train", capped_train)
    monkeypatch.setattr(Trainer, "simulate_batch", spy)
    torch.manual_seed(0)
    main_run.main(["main_run.py", "train", "one_store_lost", "vanilla_one_store"])
    assert paths and set(paths) == {"fused"}
    n_train = len(paths)
    main_run.main(["main_run.py", "test", "one_store_lost", "vanilla_one_store"])
    assert set(paths[n_train:]) == {"fused"} and len(paths) == n_train + 1  # one batch of 32768 scenarios x 5000 periods
    out = capsys.readouterr().out
    loss = float(out.strip().splitlines()[-1].split(":")[1])
    assert "Average per-period test loss" in out and np.isfinite(loss) and loss > 0


def test_two_rank_trainer_matches_single_process(tmp_path):
    """`Trainer.train` under two ranks (both on cuda:0, gloo collectives; torchrun-style environment): rank 0's batch
    permutation and initial weights are broadcast, every rank runs the fused kernels on its shard of each batch, one
    gradient all-reduce per batch - the trained parameters and the loss history must equal a single-process run
    started from rank 0's seed (VERDICT r1 "multi-GPU through the product API")."""
    import socket
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ddp_trainer_worker.py")
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    base = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    one = str(tmp_path / "one.npz")
    subprocess.run([sys.executable, worker, one], check=True, env=base, timeout=600)
    two = str(tmp_path / "two.npz")
    procs = []
    for r in range(2):
        env = dict(base, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   HDPO_DIST_BACKEND="gloo")
        procs.append(subprocess.Popen([sys.executable, worker, two], env=env))
    for pr in procs:
        assert pr.wait(timeout=600) == 0
    a, b = np.load(one), np.load(two)
    np.testing.assert_allclose(b["train"], a["train"], rtol=2e-5)
    np.testing.assert_allclose(b["dev"], a["dev"], rtol=2e-5)
    for k in a.files:
        if k.startswith("net_"):
            assert np.abs(b[k] - a[k]).max() <= 2e-4 * max(1.0, np.abs(a[k]).max()), k


def test_fused_adam_kernel_matches_torch_adam():
    """hdpo_adam_step (one launch over the flat vector) against torch.optim.Adam over 25 steps, incl. weight decay."""
    from neural_inventory_control_b200 import _capi as K, _lib
    from neural_inventory_control_b200.engine import current_stream_ptr
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    for wd in (0.0, 0.01):
        p_ref = torch.randn(10007, generator=g, device=dev).requires_grad_(True)
        opt = torch.optim.Adam([p_ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        p = p_ref.detach().clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for step in range(1, 26):
            grad = torch.randn(p.shape, generator=g, device=dev) * (0.1 + 0.05 * step)
            p_ref.grad = grad.clone()
            opt.step()
            rc = lib.hdpo_adam_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 3e-3, 0.9, 0.999,
                                    1e-8, wd, step, current_stream_ptr(dev))
            K.check(lib, rc, "hdpo_adam_step")
        torch.cuda.synchronize()
        assert float((p - p_ref.detach()).abs().max()) <= 2e-6 * float(p_ref.detach().abs().max())
        st = opt.state[p_ref]
        assert float((m - st["exp_avg"]).abs().max()) <= 1e-6 * float(st["exp_avg"].abs().max())
        assert float((v - st["exp_avg_sq"]).abs().max()) <= 1e-6 * float(st["exp_avg_sq"].abs().max())


def test_fused_training_step_matches_reference_control_flow(monkeypatch):
    """`Trainer.train` with the fused step (flat vectors, adjoint straight into the flat gradient, hdpo_adam_step, no
    per-batch host read) against the reference's control flow (autograd node + torch.optim.Adam.step) on the same
    seeds: loss histories and trained weights agree; the optimizer's state_dict stays checkpoint-compatible."""
    from torch.utils.data import DataLoader
    from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    from neural_inventory_control_b200.trainer import Trainer
    dev = "cuda:0"
    results = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("HDPO_FUSED_STEP", mode)
        s = copy.deepcopy(_cfg("settings", "one_warehouse_lost_demand"))
        p = copy.deepcopy(_cfg("policies_and_hyperparams", "vanilla_warehouse"))
        p["nn_params"]["neurons_per_hidden_layer"]["master"] = [64, 64]
        p["nn_params"]["gradient_clipping_norm_value"] = 5.0
        obs_params = defaultdict(lambda: None, s["observation_params"])
        pbd = s["params_by_dataset"]
        pbd["train"].update(n_samples=600, batch_size=256)
        pbd["dev"].update(n_samples=128, batch_size=128)
        common = (s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"])
        periods = max(pbd["train"]["periods"], pbd["dev"]["periods"])
        sc = Scenario(periods, *common, 600 + 128, obs_params, copy.deepcopy(s["seeds"]))
        train, devset = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True, sample_index_for_split=128)
        loaders = {"train": DataLoader(train, batch_size=256, shuffle=True),
                   "dev": DataLoader(devset, batch_size=128, shuffle=False)}
        torch.manual_seed(0)
        model = NeuralNetworkCreator().create_neural_network(sc, p["nn_params"], device=dev)
        opt = torch.optim.Adam(model.parameters(), lr=p["optimizer_params"]["learning_rate"])
        tr, sim = Trainer(device=dev), Simulator(device=dev)
        tp = p["trainer_params"]
        tp.update(epochs=5, do_dev_every_n_epochs=2, print_results_every_n_epochs=100, save_model=False)
        tr.train(5, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs_params, pbd, tp)
        assert tr.last_path == "fused"
        sd = opt.state_dict()
        assert len(sd["state"]) == len(list(model.parameters())) and all("exp_avg_sq" in v for v in sd["state"].values())
        assert int(next(iter(sd["state"].values()))["step"]) == 5 * 3
        results[mode] = (np.array(tr.all_train_losses), np.array(tr.all_dev_losses),
                         {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()})
    a, b = results["1"], results["0"]
    np.testing.assert_allclose(a[0], b[0], rtol=2e-5)
    np.testing.assert_allclose(a[1], b[1], rtol=2e-5)
    for k in a[2]:
        assert np.abs(a[2][k] - b[2][k]).max() <= 1e-4 * max(1.0, np.abs(b[2][k]).max()), k
