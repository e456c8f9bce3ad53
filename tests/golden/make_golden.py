#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

The reference (MatiasAlvo/Neural_inventory_control, mounted read-only at /root/reference) ships no
tests and no golden vectors, so the pins for this repo's oracle are produced
here: the reference's own `Scenario`, `NeuralNetworkCreator`, `Simulator` and
`Trainer.simulate_batch` are imported as they are (two inert import stubs for the absent
`gymnasium` / `matplotlib` live in oracle/refstubs/) and executed on CPU in fp32 and fp64.

Run (in the build container only - /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

What is stored per rollout case (np.savez_compressed, float32 unless noted):
    data/<key>            every tensor of Scenario.get_data() for the first B samples
    param/<state_dict key> policy weights handed to BOTH sides by the parity tests
    meta                  json string: problem_params, T, ignore_periods, policy name, wub, layer info
    ref/reward_tb         [T,B] per-period per-scenario cost (the `reward` Simulator.step returns)
    ref/total, ref/report the two scalars Trainer.simulate_batch returns
    ref/grad/<key>        d(total/(B*T*S))/d(param), reference autograd, fp32
    ref64/...             same quantities from the float64 run of the reference (ground truth)
    ref/final/<obs key>   final inventories (state after T periods)
    ref/action0/<key>     the action dict of period 0

and per single-step case (file step_*.npz): inputs, action, outputs and autograd gradients of one
`Simulator.step` under random upstream adjoints, including exact-zero allocations and a lead-time-0
"not connected" pair.
"""
import copy
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HDPO_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

# The shipped quantile-forecaster checkpoint was written from a CUDA process and the reference loads it with a bare
# torch.load(path): on this CPU-only container that needs a default map_location. The wrapper lives HERE (generator
# side); the reference's sources stay untouched.
_torch_load = torch.load


def _cpu_load(f, *a, **k):
    k.setdefault("map_location", "cpu")
    k.setdefault("weights_only", False)
    return _torch_load(f, *a, **k)


torch.load = _cpu_load

import trainer as ref_trainer  # noqa: E402  (the reference's trainer.py)
from collections import defaultdict  # noqa: E402

torch.set_num_threads(8)


class CapturingLoss(ref_trainer.PolicyLoss):
    """PolicyLoss that also records the per-scenario reward handed to it (trainer.py:206)."""

    def __init__(self):
        super().__init__()
        self.rewards = []
        self.actions = []

    def forward(self, observation, action, reward):
        self.rewards.append(reward.detach().clone())
        if len(self.actions) < 1:
            self.actions.append({k: v.detach().clone() for k, v in action.items()})
        return reward.sum()


def load_cfg(setting, policy):
    with open(f"{REF}/config_files/settings/{setting}.yml") as f:
        s = yaml.safe_load(f)
    with open(f"{REF}/config_files/policies_and_hyperparams/{policy}.yml") as f:
        p = yaml.safe_load(f)
    return s, p


def build_case(setting, policy, B, T, T_total, setting_patch=None, nn_patch=None, torch_seed=0,
               state_dict=None):
    s, p = load_cfg(setting, policy)
    if setting_patch:
        setting_patch(s)
    if nn_patch:
        nn_patch(p["nn_params"])
    obs_params = defaultdict(lambda: None, s["observation_params"])
    seeds = copy.deepcopy(s["seeds"])
    scenario = ref_trainer.Scenario(T_total, s["problem_params"], s["store_params"], s["warehouse_params"],
                                    s["echelon_params"], B, obs_params, seeds)
    data = {k: v.clone().contiguous() for k, v in scenario.get_data().items()}
    torch.manual_seed(torch_seed)
    model = ref_trainer.NeuralNetworkCreator().create_neural_network(scenario, p["nn_params"], device="cpu")
    return s, p, obs_params, scenario, data, model


def run_reference(s, obs_params, data, model, T, ignore, dtype, backward=True):
    """Trainer.simulate_batch + backward, exactly as trainer.py:163-173 does for one batch."""
    sim = ref_trainer.Simulator(device="cpu")
    tr = ref_trainer.Trainer(device="cpu")
    loss = CapturingLoss()
    d = {k: v.clone().to(dtype) for k, v in data.items()}
    for prm in model.parameters():
        prm.grad = None
    total, report = tr.simulate_batch(loss, sim, model, T, s["problem_params"], d, obs_params, ignore, False)
    B = len(d["demands"])
    mean_loss = total / (B * T * s["problem_params"]["n_stores"])
    if backward:
        mean_loss.backward()
    out = {
        "reward_tb": torch.stack(loss.rewards, 0).numpy(),
        "total": np.array(total.item()),
        "report": np.array(report.item()),
    }
    for k, v in model.named_parameters():
        if backward and v.grad is not None:
            out[f"grad/{k}"] = v.grad.detach().clone().numpy()
    for k in ("store_inventories", "warehouse_inventories", "echelon_inventories"):
        if k in sim.observation:
            out[f"final/{k}"] = sim.observation[k].detach().numpy()
    for k, v in loss.actions[0].items():
        out[f"action0/{k}"] = v.numpy()
    return out


ONLY = None  # --only a,b: regenerate just these cases


def rollout_case(name, setting, policy, B, T, T_total, ignore=30, setting_patch=None, nn_patch=None,
                 torch_seed=0, state_dict_path=None, perturb=0.0, compact=False, kind="rollout", backward=True):
    if ONLY is not None and name not in ONLY:
        return
    if kind == "realdata":
        os.chdir(REF)  # the real-data settings name their demand / calendar files relative to the reference root
    s, p, obs_params, scenario, data, model = build_case(setting, policy, B, T, T_total, setting_patch, nn_patch,
                                                         torch_seed)
    # materialise LazyLinear layers with one throw-away forward
    sim = ref_trainer.Simulator(device="cpu")
    obs, _ = sim.reset(T, s["problem_params"], {k: v.clone() for k, v in data.items()}, obs_params)
    o = dict(obs)
    o["internal_data"] = sim._internal_data
    with torch.no_grad():
        model(o)
    if state_dict_path is not None:
        ck = torch.load(state_dict_path, map_location="cpu", weights_only=False)
        model.load_state_dict(ck["model_state_dict"])
    if perturb:
        # move weights away from the default init so that kinks / saturation are exercised
        g = torch.Generator().manual_seed(1234)
        with torch.no_grad():
            for prm in model.parameters():
                prm.add_(perturb * torch.randn(prm.shape, generator=g))
    arrays = {}
    for k, v in data.items():
        arrays[f"data/{k}"] = v.numpy()
    for k, v in model.state_dict().items():
        if isinstance(v, torch.nn.parameter.UninitializedTensorMixin):
            continue  # benchmark policies that never call their (lazy) net
        arrays[f"param/{k}"] = v.detach().clone().numpy()
    if hasattr(model, "fixed_nets"):  # frozen quantile forecaster: not in the policy's state_dict
        for k, v in model.fixed_nets["quantile_forecaster"].state_dict().items():
            arrays[f"aux/forecaster/{k}"] = v.detach().clone().numpy()
    r32 = run_reference(s, obs_params, data, model, T, ignore, torch.float32, backward)
    for k, v in r32.items():
        arrays[f"ref/{k}"] = v
    wub = model.warehouse_upper_bound
    model64 = copy.deepcopy(model).double()
    if torch.is_tensor(wub):
        model64.warehouse_upper_bound = wub.double()
    if hasattr(model64, "fixed_nets"):  # deepcopy(...).double() does not reach the plain-dict forecaster
        model64.fixed_nets["quantile_forecaster"].double()
    if kind == "realdata":
        # policies that build constants with torch.tensor([0.0]) (FixedQuantile) need the default dtype to follow
        torch.set_default_dtype(torch.float64)
    try:
        r64 = run_reference(s, obs_params, data, model64, T, ignore, torch.float64, backward)
    finally:
        torch.set_default_dtype(torch.float32)
    for k, v in r64.items():
        if k.startswith("grad/") or k in ("reward_tb", "total", "report"):
            # the 512-wide cases keep the float64 gradient as float32 (7 significant digits of the ground truth are
            # plenty for a 1e-5 bar; halves the fixture)
            arrays[f"ref64/{k}"] = v.astype(np.float32) if (compact and k.startswith("grad/")) else v
    pp = {k: v for k, v in scenario.problem_params.items()}
    meta = {
        "setting": setting, "policy": policy, "nn_name": p["nn_params"]["name"],
        "problem_params": pp, "T": T, "ignore_periods": ignore, "B": B,
        "period_shift": s["observation_params"]["demand"]["period_shift"],
        "warehouse_upper_bound": (float(wub[0]) if torch.is_tensor(wub) else float(wub)),
        "transshipment": bool(p["nn_params"].get("transshipment", False)),
        "neurons_per_hidden_layer": p["nn_params"]["neurons_per_hidden_layer"],
        "inner_layer_activations": p["nn_params"]["inner_layer_activations"],
        "output_layer_activation": p["nn_params"]["output_layer_activation"],
        "demands_sha256_12": hashlib.sha256(data["demands"].numpy().tobytes()).hexdigest()[:12],
        "torch": torch.__version__, "numpy": np.__version__,
    }
    arrays["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, f"{kind}_{name}.npz")
    np.savez_compressed(path, **arrays)
    gn = float(np.sqrt(sum((v.astype(np.float64) ** 2).sum() for k, v in r32.items() if k.startswith("grad/"))))
    print(f"{name:34s} B={B:4d} total={r32['total']:.6e} report={r32['report']:.6e} |grad|={gn:.6e} "
          f"sha={meta['demands_sha256_12']} -> {os.path.getsize(path) / 1024:.0f} KiB")


def step_case(name, B, S, W, E, L, Lw, Le, lost, profit, edge_cost, seed, stray=False):
    """One reference `Simulator.step` with random state / action / upstream adjoints (environment.py:110-169)."""
    if ONLY is not None and f"step_{name}" not in ONLY:
        return
    g = torch.Generator().manual_seed(seed)
    Wc = max(W, 1)

    def rnd(*shape, lo=0.0, hi=1.0):
        return lo + (hi - lo) * torch.rand(*shape, generator=g)

    problem_params = {"n_stores": S, "n_warehouses": W, "n_extra_echelons": E, "lost_demand": lost,
                      "maximize_profit": profit}
    T_total = 3
    data = {
        "demands": rnd(B, S, T_total, lo=0, hi=8),
        "initial_inventories": rnd(B, S, L, lo=-2 if not lost else 0, hi=6),
        "holding_costs": rnd(B, S, lo=0.5, hi=1.5),
        "underage_costs": rnd(B, S, lo=4, hi=12),
        "lead_times": torch.randint(1, L + 1, (B, S, Wc), generator=g).float(),
    }
    if W > 0:
        data["initial_warehouse_inventories"] = rnd(B, W, Lw, lo=0, hi=30)
        data["warehouse_lead_times"] = torch.randint(1, Lw + 1, (B, W), generator=g).float()
        data["warehouse_holding_costs"] = rnd(B, W, lo=0.1, hi=0.6)
        if edge_cost:
            data["warehouse_edge_costs"] = rnd(B, W, lo=0.2, hi=1.5)
    if E > 0:
        data["initial_echelon_inventories"] = rnd(B, E, Le, lo=0, hi=30)
        data["echelon_lead_times"] = torch.randint(1, Le + 1, (B, E), generator=g).float()
        data["echelon_holding_costs"] = rnd(B, E, lo=0.05, hi=0.3)
    action = {"stores": rnd(B, S, Wc, lo=0, hi=5)}
    # exact zeros (no pipeline gradient, environment.py:426-432) and a lead-time-0 "not connected" pair
    zero_mask = torch.rand(B, S, Wc, generator=g) < 0.25
    action["stores"][zero_mask] = 0.0
    if W > 1:
        data["lead_times"][:, 0, 0] = 0.0
        action["stores"][:, 0, 0] = 0.0
    if stray:
        # NON-zero orders on lead-time-0 pairs (what the shipped GNN policy does on many_warehouses_lost_demand): the
        # reference's flat put lands them in the previous node's last slot; (b=0, s=0) wraps to the tensor's last element
        data["lead_times"][:, 0, 0] = 0.0
        action["stores"][:, 0, 0] = rnd(B, lo=0.5, hi=3.0)
        data["lead_times"][:, 2, 1] = 0.0
        action["stores"][:, 2, 1] = rnd(B, lo=0.5, hi=3.0)
    # some on-hand inventories exactly equal to demand -> clip kink at 0 (grad 1 at x==0)
    data["initial_inventories"][0, :, 0] = data["demands"][0, :, 1]
    if W > 0:
        action["warehouses"] = rnd(B, W, 1, lo=0, hi=9)
        action["warehouses"][torch.rand(B, W, 1, generator=g) < 0.2] = 0.0
    if E > 0:
        action["echelons"] = rnd(B, E, 1, lo=0, hi=9)
        action["echelons"][torch.rand(B, E, 1, generator=g) < 0.2] = 0.0
    obs_params = defaultdict(lambda: None, {
        "include_warehouse_inventory": W > 0,
        "include_static_features": {"holding_costs": True, "underage_costs": True, "lead_times": True},
        "demand": {"past_periods": 0, "period_shift": 1},
    })
    arrays = {}
    for dtype, tag in ((torch.float32, "ref"), (torch.float64, "ref64")):
        d = {k: v.clone().to(dtype) for k, v in data.items()}
        a = {k: v.clone().to(dtype).requires_grad_(True) for k, v in action.items()}
        inv_keys = ["initial_inventories"] + (["initial_warehouse_inventories"] if W > 0 else []) + \
                   (["initial_echelon_inventories"] if E > 0 else [])
        for k in inv_keys:
            d[k].requires_grad_(True)
        sim = ref_trainer.Simulator(device="cpu")
        obs, _ = sim.reset(2, problem_params, d, obs_params)
        shifts = {k: sim._internal_data[k].numpy().copy() for k in sim._internal_data if k.endswith("allocation_shift")}
        obs, reward, terminated, _, _ = sim.step(a)
        gg = torch.Generator().manual_seed(seed + 1)
        r_bar = torch.rand(B, generator=gg).to(dtype)
        loss = (reward * r_bar).sum()
        ups = {}
        for k in ("store_inventories", "warehouse_inventories", "echelon_inventories"):
            if k in obs and torch.is_tensor(obs[k]) and obs[k].requires_grad:
                ups[k] = torch.randn(obs[k].shape, generator=gg).to(dtype)
                loss = loss + (obs[k] * ups[k]).sum()
        loss.backward()
        if tag == "ref":
            for k, v in data.items():
                arrays[f"data/{k}"] = v.numpy()
            for k, v in action.items():
                arrays[f"action/{k}"] = v.numpy()
            arrays["up/reward"] = r_bar.numpy()
            for k, v in ups.items():
                arrays[f"up/{k}"] = v.numpy()
            for k, v in shifts.items():
                arrays[f"ref/{k}"] = v  # int64, bit-exact pins (environment.py:77-101)
        arrays[f"{tag}/reward"] = reward.detach().numpy()
        for k in ups:
            arrays[f"{tag}/new/{k}"] = obs[k].detach().numpy()
        for k in inv_keys:
            arrays[f"{tag}/grad/{k}"] = d[k].grad.numpy()
        for k, v in a.items():
            arrays[f"{tag}/grad/action_{k}"] = v.grad.numpy()
    meta = {"problem_params": problem_params, "period": 0, "period_shift": 1, "B": B}
    arrays["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, f"step_{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"step_{name:29s} reward[0]={arrays['ref/reward'][0]:.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def synthetic_many_warehouses(S, W, seed=7):
    """cfg 5 shape: Bernoulli(0.7) adjacency, every store connected, leads in [1,7)."""
    rng = np.random.RandomState(seed)
    adj = (rng.rand(W, S) < 0.7).astype(int)
    for s_ in range(S):
        if adj[:, s_].sum() == 0:
            adj[rng.randint(W), s_] = 1
    lead = rng.randint(1, 7, size=(S, W)) * adj.T
    return adj.tolist(), lead.tolist()


def main():
    def lead20(s):
        s["store_params"]["lead_time"]["value"] = 20

    def stores50(s):
        s["problem_params"]["n_stores"] = 50

    def many3x50(s):
        adj, lead = synthetic_many_warehouses(50, 3)
        s["problem_params"]["n_stores"] = 50
        s["problem_params"]["n_warehouses"] = 3
        s["problem_params"]["warehouse_store_adjacency"] = adj
        s["store_params"]["lead_time"]["value"] = lead
        s["warehouse_params"]["holding_cost"] = [0.3, 0.4, 0.2]
        s["warehouse_params"]["edge_cost"] = [0.5, 1.5, 0.7]

    def hidden(widths):
        def f(nn):
            nn["neurons_per_hidden_layer"]["master"] = widths
        return f

    rollout_case("one_store_lost", "one_store_lost", "vanilla_one_store", B=64, T=50, T_total=100)
    rollout_case("one_store_lost_trained", "one_store_lost", "vanilla_one_store", B=64, T=50, T_total=100,
                 state_dict_path=f"{REF}/saved_models/2024_04_23/vanilla_one_store/1713902211.pt")
    rollout_case("one_store_backlogged", "one_store_backlogged", "vanilla_one_store", B=64, T=50, T_total=60)
    rollout_case("one_store_backlogged_lead20", "one_store_backlogged", "vanilla_one_store", B=64, T=50, T_total=50,
                 setting_patch=lead20, perturb=0.05)
    rollout_case("serial_system", "serial_system", "vanilla_serial", B=64, T=50, T_total=100)
    rollout_case("serial_system_perturbed", "serial_system", "vanilla_serial", B=64, T=50, T_total=50, perturb=0.2)
    rollout_case("one_warehouse_s5", "one_warehouse_lost_demand", "vanilla_warehouse", B=32, T=50, T_total=60,
                 nn_patch=hidden([64, 64, 64]))
    rollout_case("one_warehouse_s50", "one_warehouse_lost_demand", "vanilla_warehouse", B=16, T=50, T_total=50,
                 setting_patch=stores50, nn_patch=hidden([64, 48]), perturb=0.02)
    rollout_case("many_warehouses_2x10", "many_warehouses_lost_demand", "vanilla_warehouse", B=32, T=50, T_total=60,
                 nn_patch=hidden([64, 64, 64]))
    rollout_case("many_warehouses_3x50", "many_warehouses_lost_demand", "vanilla_warehouse", B=16, T=50, T_total=50,
                 setting_patch=many3x50, nn_patch=hidden([64, 48]), perturb=0.02)
    # the widths bench.py measures (BASELINE cfg 4 / cfg 5 with the shipped vanilla_warehouse.yml: three 512-wide layers)
    rollout_case("one_warehouse_s50_w512", "one_warehouse_lost_demand", "vanilla_warehouse", B=16, T=50, T_total=50,
                 setting_patch=stores50, compact=True)
    rollout_case("many_warehouses_3x50_w512", "many_warehouses_lost_demand", "vanilla_warehouse", B=16, T=50,
                 T_total=50, setting_patch=many3x50, compact=True)

    # the reference's shipped weight-shared policy (GNN) on the generic per-step path: kind "gnn" keeps these fixtures
    # out of the fused-kernel / numpy-oracle test matrices (the oracle restates the fused policies only)
    rollout_case("one_warehouse", "one_warehouse_lost_demand", "gnn", B=32, T=50, T_total=60, kind="gnn")
    rollout_case("many_warehouses", "many_warehouses_lost_demand", "gnn", B=32, T=50, T_total=60, kind="gnn")
    def with_moments(s):  # the shipped transshipment setting omits the mean / std features the GNN reads (KeyError)
        s["observation_params"]["include_static_features"].update(mean=True, std=True)

    rollout_case("transshipment", "transshipment_backlogged", "gnn_transshipment", B=32, T=50, T_total=60, kind="gnn",
                 setting_patch=with_moments)

    # real-data observation features (past demands, days from Christmas, period shift, profit objective) with the
    # policies that consume them: DataDrivenNet and the clairvoyant benchmark on the shipped 21-store / 3-warehouse
    # Favorita setting; the quantile policies on one-store samples cut from the same file (the reference does not ship
    # data_files/favorita/weekly_sales.pt, so the one-store setting gets its demand from a reshaped copy under /tmp)
    def real_one_store(s):
        tmp = "/tmp/hdpo_weekly_one_store.pt"
        w = _torch_load(f"{REF}/data_files/favorita_21_stores/weekly_sales.pt")
        torch.save(w.reshape(-1, 1, w.shape[2]).contiguous(), tmp)
        s["store_params"]["demand"]["file_location"] = tmp

    rollout_case("many_warehouses_data_driven", "many_warehouses_real_data_lost_demand", "data_driven_net", B=16, T=50,
                 T_total=171, ignore=16, kind="realdata")
    rollout_case("many_warehouses_just_in_time", "many_warehouses_real_data_lost_demand", "just_in_time", B=16, T=50,
                 T_total=171, ignore=16, kind="realdata", backward=False)
    for pol in ("transformed_nv", "fixed_quantile"):
        rollout_case(f"one_store_{pol}", "one_store_real_data_lost_demand", pol, B=48, T=50, T_total=171, ignore=16,
                     setting_patch=real_one_store, kind="realdata")
    for pol in ("quantile_nv", "returns_nv", "just_in_time"):
        rollout_case(f"one_store_{pol}", "one_store_real_data_lost_demand", pol, B=48, T=50, T_total=171, ignore=16,
                     setting_patch=real_one_store, kind="realdata", backward=False)

    step_case("one_store_lost", B=16, S=1, W=0, E=0, L=4, Lw=0, Le=0, lost=True, profit=False, edge_cost=False, seed=1)
    step_case("one_store_backlog_profit", B=16, S=1, W=0, E=0, L=7, Lw=0, Le=0, lost=False, profit=True,
              edge_cost=False, seed=2)
    step_case("serial", B=16, S=1, W=1, E=2, L=4, Lw=3, Le=4, lost=False, profit=False, edge_cost=False, seed=3)
    step_case("one_warehouse", B=8, S=7, W=1, E=0, L=3, Lw=3, Le=0, lost=True, profit=False, edge_cost=False, seed=4)
    step_case("many_warehouses", B=8, S=6, W=3, E=0, L=6, Lw=3, Le=0, lost=True, profit=False, edge_cost=True,
              seed=5)
    step_case("many_warehouses_stray", B=8, S=6, W=3, E=0, L=6, Lw=3, Le=0, lost=True, profit=False, edge_cost=True,
              seed=6, stray=True)


if __name__ == "__main__":
    if "--only" in sys.argv:
        ONLY = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    main()
