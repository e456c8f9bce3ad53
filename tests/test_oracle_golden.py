"""Pin the numpy oracle (oracle/hdpo_oracle.py) against outputs of the unmodified reference.

The fixtures were produced by tests/golden/make_golden.py running the reference's own Scenario /
Simulator / Trainer.simulate_batch / autograd on CPU (fp32 and fp64). Tolerances:
  * integer index tables (allocation_shift): bit-exact
  * single-step outputs and gradients: fp64 run 1e-12, fp32 run 2e-6 relative
  * rollout per-scenario per-period costs: fp64 1e-9; fp32 1e-5 relative (north-star bar)
  * rollout parameter gradients: fp64 1e-8 relative L2; fp32 within max(1e-5, 3x the reference's own
    fp32-vs-fp64 error) (the wide nets' fp32 noise floor is ~4e-5)
"""
import numpy as np
import pytest

from oracle import hdpo_oracle as O
import golden_util as G


@pytest.mark.parametrize("name", G.step_cases())
@pytest.mark.parametrize("tag,dtype,tol", [("ref64", np.float64, 1e-12), ("ref", np.float32, 2e-6)])
def test_single_step_matches_reference(name, tag, dtype, tol):
    meta, g = G.load("step", name)
    pb = G.problem_from_meta(meta)
    data = G.cast(g["data"], dtype)
    action = G.cast(g["action"], dtype)
    up = G.cast(g["up"], dtype)
    ref = g[tag]
    state = O._initial_state(pb, data)
    B = data["demands"].shape[0]
    # integer tables: bit-exact (environment.py:77-101)
    if tag == "ref":
        for key, inv in (("allocation_shift", "store"), ("warehouse_allocation_shift", "wh"),
                         ("echelon_allocation_shift", "ech")):
            if key in ref:
                n, L = state[inv].shape[1:]
                got = O.allocation_shift(B, n, L)
                assert got.dtype == np.int64 and np.array_equal(got, ref[key])
    new, reward, esaved = O.env_step(pb, state, action, data, meta["period"])
    assert G.rel_l2(reward, ref["reward"]) < tol
    names = {"store": "store_inventories", "wh": "warehouse_inventories", "ech": "echelon_inventories"}
    for k, v in new.items():
        np.testing.assert_allclose(v, ref[f"new/{names[k]}"], rtol=tol, atol=tol)
    # adjoint with the same upstream: loss = sum(reward*r_b) + sum(new*up); r_bar is per-scenario here
    g_new = {k: up[names[k]] for k in new}
    r_bar = up["reward"]
    # (a [B,1] r_bar broadcasts where the trainer's scalar 1/(B*T*S) would)
    g_state, g_act = O.env_step_adjoint(pb, state, action, data, esaved, g_new, r_bar[:, None])
    inv_names = {"store": "initial_inventories", "wh": "initial_warehouse_inventories",
                 "ech": "initial_echelon_inventories"}
    for k, v in g_state.items():
        np.testing.assert_allclose(v, ref[f"grad/{inv_names[k]}"], rtol=10 * tol, atol=10 * tol)
    for k, v in g_act.items():
        np.testing.assert_allclose(v, ref[f"grad/action_{k}"], rtol=10 * tol, atol=10 * tol)


@pytest.mark.parametrize("name", G.rollout_cases())
def test_rollout_fp64_matches_reference_fp64(name):
    meta, g = G.load("rollout", name)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    data = G.cast(g["data"], np.float64)
    fwd, grads = O.rollout_grad(pol, pb, data, meta["T"])
    ref = g["ref64"]
    np.testing.assert_allclose(fwd["reward_tb"], ref["reward_tb"], rtol=1e-9, atol=1e-9)
    assert abs(fwd["total"] - ref["total"]) <= 1e-10 * abs(ref["total"])
    rep = fwd["reward_tb"][meta["ignore_periods"]:].sum()
    assert abs(rep - ref["report"]) <= 1e-10 * abs(ref["report"])
    flat = O.flatten_grads(pol, grads)
    for k, v in flat.items():
        # the 512-wide fixtures store the float64 gradient rounded to float32 (half the file size): 2^-24 per element
        tol = 1e-8 if ref[f"grad/{k}"].dtype == np.float64 else 1e-7
        assert G.rel_l2(v, ref[f"grad/{k}"]) < tol, k


@pytest.mark.parametrize("name", G.rollout_cases())
def test_rollout_fp32_matches_reference_fp32(name):
    meta, g = G.load("rollout", name)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float32)
    data = G.cast(g["data"], np.float32)
    fwd, grads = O.rollout_grad(pol, pb, data, meta["T"])
    ref, ref64 = g["ref"], g["ref64"]
    assert fwd["reward_tb"].dtype == np.float32
    # per-scenario (summed over periods) costs: the north-star 1e-5 relative bar
    # gate against the float64 truth; allowance = max(1e-5, 3x the reference's own fp32 error to that truth)
    cost_b = fwd["reward_tb"].astype(np.float64).sum(0)
    ref_b = ref["reward_tb"].astype(np.float64).sum(0)
    true_b = ref64["reward_tb"].sum(0)
    cfloor = np.abs(ref_b / true_b - 1).max()
    assert np.abs(cost_b / true_b - 1).max() <= max(1e-5, 3 * cfloor), (np.abs(cost_b / true_b - 1).max(), cfloor)
    assert abs(float(fwd["total"]) - float(ref64["total"])) <= max(1e-5, 3 * cfloor) * abs(float(ref64["total"]))
    flat = O.flatten_grads(pol, grads)
    mine = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in sorted(flat)])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in sorted(flat)])
    floor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, 3 * floor), (G.rel_l2(mine, r64), floor)


def test_final_state_and_first_action():
    for name in G.rollout_cases():
        meta, g = G.load("rollout", name)
        pb = G.problem_from_meta(meta)
        pol = G.policy_from_golden(meta, g["param"], np.float64)
        data = G.cast(g["data"], np.float64)
        state = O._initial_state(pb, data)
        action, _ = O.policy_forward(pol, pb, state, data)
        for k, v in action.items():
            np.testing.assert_allclose(v, g["ref"][f"action0/{k}"], rtol=2e-4, atol=2e-5)  # fp32 reference action
        fwd = O.rollout_forward(pol, pb, data, meta["T"])
        names = {"store": "store_inventories", "wh": "warehouse_inventories", "ech": "echelon_inventories"}
        for k, v in fwd["final"].items():
            ref = g["ref"][f"final/{names[k]}"]
            assert v.shape == ref.shape
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(v - ref).max() <= 2e-3 * scale, (name, k)  # fp32 reference vs fp64 oracle after T periods


def test_discrete_allocation_rounds_half_to_even():
    meta, g = G.load("rollout", "one_store_lost")
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float32)
    data = G.cast(g["data"], np.float32)
    fwd = O.rollout_forward(pol, pb, data, 5, discrete=True, keep_tape=True)
    for _, action, _, _ in fwd["tape"]:
        assert np.array_equal(action["stores"], np.rint(action["stores"]))


# ---- the PyTorch-eager CPU port used as the reported CPU baseline (oracle/torch_port.py) ----
def _torch_policy(meta, params, dtype):
    import torch
    idxs = sorted({int(k.split(".")[2]) for k in params if k.endswith(".weight")})
    layers = [(torch.tensor(params[f"net.master.{i}.weight"], dtype=dtype, requires_grad=True),
               torch.tensor(params[f"net.master.{i}.bias"], dtype=dtype, requires_grad=True)) for i in idxs]
    adj = meta["problem_params"].get("warehouse_store_adjacency")
    return {"arch": meta["nn_name"], "layers": layers, "hidden_act": meta["inner_layer_activations"]["master"],
            "out_act": meta["output_layer_activation"]["master"],
            "wub": torch.tensor([meta["warehouse_upper_bound"]], dtype=dtype),
            "adjacency": None if adj is None else torch.tensor(adj), "transshipment": meta.get("transshipment", False)}


@pytest.mark.parametrize("name", G.rollout_cases())
def test_torch_port_matches_reference_fp64(name):
    import torch
    from oracle import torch_port as TP
    meta, g = G.load("rollout", name)
    pol = _torch_policy(meta, g["param"], torch.float64)
    pb = dict(meta["problem_params"], period_shift=meta.get("period_shift", 0))
    data = {k: torch.tensor(v, dtype=torch.float64) for k, v in g["data"].items()}
    total, report, grads = TP.train_step(pol, pb, data, meta["T"], meta["ignore_periods"])
    ref = g["ref64"]
    assert abs(total - ref["total"]) <= 1e-10 * abs(ref["total"])
    assert abs(report - ref["report"]) <= 1e-10 * abs(ref["report"])
    idxs = sorted({int(k.split(".")[2]) for k in g["param"] if k.endswith(".weight")})
    names = [f"net.master.{i}.{wb}" for i in idxs for wb in ("weight", "bias")]
    for n, gr in zip(names, grads):
        tol = 1e-8 if ref[f"grad/{n}"].dtype == np.float64 else 1e-7  # 512-wide fixtures: float32-rounded float64 truth
        assert G.rel_l2(gr.numpy(), ref[f"grad/{n}"]) < tol, n


def test_symmetry_aware_oracle_adjoint_matches_torch_autograd():
    """SymmetryAware has no executable reference (recovered from stale bytecode), so the explicit
    adjoint of the oracle restatement is cross-checked against torch autograd through this repo's torch policy
    class + the pinned torch port of the simulator, in float64 (parity 'restatement vs restatement')."""
    import copy
    import os
    import torch
    import yaml
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    from oracle import torch_port as TP
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    meta, g = G.load("rollout", "one_warehouse_s5")
    p = yaml.safe_load(open(os.path.join(root, "config_files/policies_and_hyperparams/symmetry_aware.yml")))
    p["nn_params"]["neurons_per_hidden_layer"] = {"context": [24], "store": [16, 16], "warehouse": [8, 8]}
    p["nn_params"]["output_sizes"]["context"] = 12

    class Scen:
        problem_params = meta["problem_params"]
        store_params = {"demand": {"mean": [meta["warehouse_upper_bound"] / 4]}}
    torch.manual_seed(3)
    model = NeuralNetworkCreator().create_neural_network(Scen(), p["nn_params"], device="cpu")
    data = {k: torch.tensor(v[:16]).double() for k, v in g["data"].items()}

    def obs_of(state):
        return {"store_inventories": state["store"], "warehouse_inventories": state["wh"], "mean": data["mean"],
                "std": data["std"], "underage_costs": data["underage_costs"], "lead_times": data["lead_times"]}
    state = {"store": data["initial_inventories"], "wh": data["initial_warehouse_inventories"]}
    model({k: v.float() for k, v in obs_of(state).items()})  # materialise the lazy layers
    model = model.double()
    model.warehouse_upper_bound = model.warehouse_upper_bound.double()
    pb = dict(meta["problem_params"], period_shift=0)
    T, total = 10, 0
    for t in range(T):
        state, r = TP.env_step(pb, state, model(obs_of(state)), data, t)
        total = total + r.sum()
    (total / (16 * T * pb["n_stores"])).backward()
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    a = p["nn_params"]
    nets = {m: O.mlp_from_state_dict(sd, m, a["inner_layer_activations"][m], a["output_layer_activation"][m])
            for m in ("context", "store", "warehouse")}
    pol = O.Policy("symmetry_aware", nets, float(model.warehouse_upper_bound[0]), prop_eps=1e-15)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), {k: v.numpy() for k, v in data.items()}, T)
    assert abs(float(total.detach()) / fwd["total"] - 1) < 1e-12
    flat = O.flatten_grads(pol, grads)
    for k, v in model.named_parameters():
        assert G.rel_l2(v.grad.numpy(), flat[k]) < 1e-10, k


def test_torch_port_symmetry_aware_matches_oracle():
    """The PyTorch-eager port of the SymmetryAware policy (what `bench.py --impl reference` / `cpu_baseline` time for the
    symmetry-aware workload) against the numpy oracle: total cost and all gradients, float64."""
    import torch
    from oracle import torch_port as TP
    meta, g = G.load("rollout", "one_warehouse_s5")
    rng = np.random.default_rng(4)
    S, L = g["data"]["initial_inventories"].shape[1:]
    Lw, C = g["data"]["initial_warehouse_inventories"].shape[2], 12
    widths = {"context": [S * L + Lw, 24, C], "store": [L + 4 + C, 16, 16, 1], "warehouse": [Lw + C, 8, 8, 1]}
    acts = {"context": ("elu", "sigmoid"), "store": ("elu", "softplus"), "warehouse": ("elu", "sigmoid")}
    sd, nets_t = {}, {}
    for m, ws in widths.items():
        layers = []
        for i in range(len(ws) - 1):
            w = rng.uniform(-1, 1, (ws[i + 1], ws[i])) / np.sqrt(ws[i])
            b = rng.uniform(-1, 1, ws[i + 1]) / np.sqrt(ws[i])
            sd[f"net.{m}.{2 * i}.weight"], sd[f"net.{m}.{2 * i}.bias"] = w, b
            layers.append((torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)))
        nets_t[m] = (layers, *acts[m])
    wub = float(meta["warehouse_upper_bound"])
    pol_t = {"arch": "symmetry_aware", "nets": nets_t, "layers": [wb for m in nets_t for wb in nets_t[m][0]],
             "wub": torch.tensor([wub], dtype=torch.float64), "prop_eps": 1e-15}
    data = {k: torch.tensor(v[:12]).double() for k, v in g["data"].items()}
    T = 8
    pb = dict(meta["problem_params"], period_shift=0)
    total, _, _ = TP.simulate(pol_t, pb, data, T)
    (total / (12 * T * pb["n_stores"])).backward()
    nets = {m: O.mlp_from_state_dict(sd, m, *acts[m]) for m in widths}
    pol = O.Policy("symmetry_aware", nets, wub, prop_eps=1e-15)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), {k: v.numpy() for k, v in data.items()}, T)
    assert abs(float(total.detach()) / fwd["total"] - 1) < 1e-12
    flat = O.flatten_grads(pol, grads)
    for m in widths:
        for i, (w, b) in enumerate(nets_t[m][0]):
            assert G.rel_l2(w.grad.numpy(), flat[f"net.{m}.{2 * i}.weight"]) < 1e-10, (m, i)
            assert G.rel_l2(b.grad.numpy(), flat[f"net.{m}.{2 * i}.bias"]) < 1e-10, (m, i)
