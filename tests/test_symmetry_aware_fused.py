"""Fused SymmetryAware rollout (trunk GEMMs + rollout_sym.cu heads) through the C ABI against the float64 oracle.

The policy class is absent from the reference snapshot's sources (recovered from stale bytecode), so the
chain of evidence is: reference simulator == oracle simulator (golden fixtures), oracle SymmetryAware adjoint == torch
autograd of the re-created policy (test_oracle_golden.py), and here kernels == oracle on the reference's exported
scenario data with seeded random weights. Bars: costs 1e-5 relative, gradients 2e-5 relative L2 (fp32 / 3xTF32).
"""
import copy

import numpy as np
import pytest

import abi_driver as D
import golden_util as G
from oracle import hdpo_oracle as O

_cache = {}


def backend(name):
    if name not in _cache:
        _cache[name] = D.EmuBackend() if name == "emu" else D.CudaBackend()
    return _cache[name]


MODES = [pytest.param("emu", "fp32", id="emu-fp32"),
         pytest.param("cuda", "fp32", id="cuda-fp32", marks=pytest.mark.gpu),
         pytest.param("cuda", "tf32x3", id="cuda-tf32x3", marks=pytest.mark.gpu)]


def sym_case(name, hidden, ctx_out, seed, hidden_act="elu", scale=1.0):
    """meta + seeded random parameters of a SymmetryAware policy on the scenario data of golden `name`."""
    meta0, g = G.load("rollout", name)
    meta = copy.deepcopy(meta0)
    data = g["data"]
    S, L = data["initial_inventories"].shape[1:]
    Lw = data["initial_warehouse_inventories"].shape[2]
    meta["nn_name"] = "symmetry_aware"
    meta["neurons_per_hidden_layer"] = hidden
    meta["inner_layer_activations"] = {m: hidden_act for m in hidden}
    meta["output_layer_activation"] = {"context": "sigmoid", "store": "softplus", "warehouse": "sigmoid"}
    meta["prop_eps"] = 1e-15
    rng = np.random.default_rng(seed)
    ins = {"context": S * L + Lw, "store": L + 4 + ctx_out, "warehouse": Lw + ctx_out}
    outs = {"context": ctx_out, "store": 1, "warehouse": 1}
    params = {}
    for m in ("context", "store", "warehouse"):
        widths = [ins[m]] + list(hidden[m]) + [outs[m]]
        for i in range(len(widths) - 1):
            bound = scale / np.sqrt(widths[i])  # torch's default Linear init range
            params[f"net.{m}.{2 * i}.weight"] = rng.uniform(-bound, bound, (widths[i + 1], widths[i])).astype(np.float32)
            params[f"net.{m}.{2 * i}.bias"] = rng.uniform(-bound, bound, widths[i + 1]).astype(np.float32)
    # make the proportional-allocation clip bind for some scenarios and not for others
    params[f"net.store.{2 * len(hidden['store'])}.bias"] += np.float32(1.0)
    return meta, params, data


def oracle_policy(meta, params):
    nets = {m: O.mlp_from_state_dict(params, m, meta["inner_layer_activations"][m],
                                     meta["output_layer_activation"][m]).astype(np.float64)
            for m in ("context", "store", "warehouse")}
    return O.Policy("symmetry_aware", nets, meta["warehouse_upper_bound"], prop_eps=meta["prop_eps"])


def check(out, fwd, grads, pol, ignore, gtol=2e-5):
    scale = np.abs(fwd["reward_tb"]).max()
    assert np.abs(out["reward_tb"] - fwd["reward_tb"]).max() <= 1e-5 * scale
    np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
    np.testing.assert_allclose(out["report_b"], fwd["reward_tb"][ignore:].sum(0), rtol=1e-5, atol=1e-5 * scale)
    for k in ("store", "wh"):
        np.testing.assert_allclose(out["final"][k], fwd["final"][k], rtol=1e-4, atol=1e-4)
    if grads is None:
        return
    flat = O.flatten_grads(pol, grads)
    assert sorted(flat) == sorted(out["grad"])
    mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    assert G.rel_l2(mine, want) <= gtol, G.rel_l2(mine, want)
    for k in flat:  # every block on its own (a misplaced block would hide in the global norm)
        assert G.rel_l2(out["grad"][k], flat[k]) <= 20 * gtol, (k, G.rel_l2(out["grad"][k], flat[k]))


@pytest.mark.parametrize("be_name,precision", MODES)
@pytest.mark.parametrize("name,n,T,ignore,hidden,ctx_out", [
    ("one_warehouse_s5", 19, 6, 2, {"context": [24], "store": [16, 16], "warehouse": [8, 8]}, 12),
    ("one_warehouse_s5", 32, 5, 0, {"context": [40, 24], "store": [32], "warehouse": [16, 16, 8]}, 20),
    ("one_warehouse_s50", 7, 4, 1, {"context": [64], "store": [32, 32], "warehouse": [16, 16]}, 32),
])
def test_symmetry_aware_fused_against_oracle(be_name, precision, name, n, T, ignore, hidden, ctx_out):
    be = backend(be_name)
    meta, params, data = sym_case(name, hidden, ctx_out, seed=11)
    data = D.slice_batch(data, n)
    out = D.rollout(be, meta, params, data, T=T, ignore=ignore, precision=precision)
    pol = oracle_policy(meta, params)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(data, np.float64), T)
    check(out, fwd, grads, pol, ignore)


@pytest.mark.parametrize("be_name,precision", MODES)
def test_symmetry_aware_fused_time_major_and_forward_only(be_name, precision):
    from neural_inventory_control_b200 import _capi as K
    be = backend(be_name)
    meta, params, data = sym_case("one_warehouse_s5", {"context": [24], "store": [16, 16], "warehouse": [8, 8]}, 12, 5)
    data = D.slice_batch(data, 9)
    a = D.rollout(be, meta, params, data, T=5, ignore=1, precision=precision)
    b = D.rollout(be, meta, params, data, T=5, ignore=1, precision=precision, demand_layout=K.DEMAND_TSB)
    np.testing.assert_array_equal(a["cost_b"], b["cost_b"])
    np.testing.assert_array_equal(a["grad_flat"], b["grad_flat"])
    c = D.rollout(be, meta, params, data, T=5, ignore=1, precision=precision, backward=False)
    np.testing.assert_array_equal(a["cost_b"], c["cost_b"])
    # discrete allocations (trainer.py:201-202): rounded store / warehouse orders, forward only
    dsc = D.rollout(be, meta, params, data, T=5, ignore=1, precision=precision, backward=False, discrete=True)
    pol = oracle_policy(meta, params)
    fwd = O.rollout_forward(pol, G.problem_from_meta(meta), G.cast(data, np.float64), 5, ignore_periods=1,
                            discrete=True)
    np.testing.assert_allclose(dsc["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)


@pytest.mark.parametrize("be_name,precision", MODES)
@pytest.mark.parametrize("variant", ["backlogged_edge_cost", "maximize_profit", "relu_tanh"])
def test_symmetry_aware_fused_problem_variants(be_name, precision, variant):
    """Simulator switches the shipped one_warehouse setting does not use: backlogged demand (no clip of the post-demand
    inventory), warehouse edge costs (environment.py:254), the profit objective with its min() tie convention
    (environment.py:190-194), other activations."""
    be = backend(be_name)
    act = "relu" if variant == "relu_tanh" else "elu"
    meta, params, data = sym_case("one_warehouse_s5", {"context": [24], "store": [16, 16], "warehouse": [8, 8]}, 12,
                                  seed=21, hidden_act=act)
    data = dict(D.slice_batch(data, 13))
    meta["problem_params"] = dict(meta["problem_params"])
    if variant == "backlogged_edge_cost":
        meta["problem_params"]["lost_demand"] = False
        data["warehouse_edge_costs"] = np.full_like(data["warehouse_holding_costs"], 0.7)
    elif variant == "maximize_profit":
        meta["problem_params"]["maximize_profit"] = True
        data["demands"] = np.round(data["demands"])           # integer demands and inventories: exact ties in min()
        data["initial_inventories"] = np.round(data["initial_inventories"])
    else:
        meta["inner_layer_activations"]["warehouse"] = "tanh"
    T, ignore = 6, 2
    out = D.rollout(be, meta, params, data, T=T, ignore=ignore, precision=precision)
    pol = oracle_policy(meta, params)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(data, np.float64), T)
    check(out, fwd, grads, pol, ignore)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_symmetry_aware_fused_default_widths_many_scenarios(precision):
    """The shipped symmetry_aware.yml widths (context 153->256->256, store 263->32->32->1, warehouse 259->16->16->1)
    on 50 stores, thousands of scenarios (several scenarios per persistent warp, many row tiles in the trunk GEMMs),
    ragged against the 128-row tiles."""
    be = backend("cuda")
    meta, params, data = sym_case("one_warehouse_s50", {"context": [256], "store": [32, 32], "warehouse": [16, 16]},
                                  256, seed=3)
    reps = 4096 // 16 + 3
    rng = np.random.default_rng(0)
    big = {k: np.concatenate([v] * reps)[:4096 + 37] for k, v in data.items()}
    big["demands"] = (big["demands"] * rng.uniform(0.5, 1.5, big["demands"].shape)).astype(np.float32)
    big["initial_inventories"] = (big["initial_inventories"]
                                  * rng.uniform(0.5, 1.5, big["initial_inventories"].shape)).astype(np.float32)
    T, ignore = 4, 1
    out = D.rollout(be, meta, params, big, T=T, ignore=ignore, precision=precision)
    pol = oracle_policy(meta, params)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(big, np.float64), T)
    check(out, fwd, grads, pol, ignore)


@pytest.mark.parametrize("be_name,precision", MODES)
def test_symmetry_aware_fused_more_than_64_stores(be_name, precision):
    """70 stores (the 50-store scenario data with 20 stores repeated): three 32-store passes in the adjoint head, a
    second two-stores-per-lane pass in the forward head, a state row that is not a multiple of the 64-float tile."""
    be = backend(be_name)
    meta0, g = G.load("rollout", "one_warehouse_s50")
    S2 = 70
    pick = np.concatenate([np.arange(50), np.arange(20)])
    data = {k: (v[:5][:, pick] if v.ndim >= 2 and v.shape[1] == 50 else v[:5]) for k, v in g["data"].items()}
    meta, params, _ = sym_case("one_warehouse_s50", {"context": [48], "store": [24, 24], "warehouse": [8]}, 16, seed=2)
    meta["problem_params"] = dict(meta["problem_params"], n_stores=S2)
    rng = np.random.default_rng(8)
    L, Lw = data["initial_inventories"].shape[2], data["initial_warehouse_inventories"].shape[2]
    w0 = params["net.context.0.weight"]
    params["net.context.0.weight"] = (rng.uniform(-1, 1, (w0.shape[0], S2 * L + Lw)) / np.sqrt(S2 * L + Lw)).astype(np.float32)
    T, ignore = 4, 1
    out = D.rollout(be, meta, params, data, T=T, ignore=ignore, precision=precision)
    pol = oracle_policy(meta, params)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(data, np.float64), T)
    check(out, fwd, grads, pol, ignore)
