"""Parity of the OPT-IN persistent one-launch sweeps of the wide path (csrc/wide_persist.cu) through the C ABI.

The default wide path is the per-period kernel chain of rollout_wide.cu; `hdpo_debug_set_wide_persist(1)` (or
HDPO_WIDE_PERSIST=1) routes tcgen05-mode VanillaWarehouse rollouts through one persistent kernel per direction
(CTA-pair GEMM tiles with dependency flags + dedicated policy-head CTAs). Same bars as the default path: the reference's
goldens, the float64 oracle on short horizons, and the default path itself on a multi-group batch.
"""
import numpy as np
import pytest

import abi_driver as D
import golden_util as G
from neural_inventory_control_b200 import _capi as K
from oracle import hdpo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def persist():
    be = D.CudaBackend()
    be.lib.hdpo_debug_set_wide_persist(1)
    try:
        yield be
    finally:
        be.lib.hdpo_debug_set_wide_persist(0)


def _oracle(meta, g, data, T):
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    fwd, grads = O.rollout_grad(pol, pb, G.cast(data, np.float64), T)
    flat = O.flatten_grads(pol, grads)
    return fwd, flat


@pytest.mark.parametrize("name,n,T,ignore", [("one_warehouse_s5", 32, 6, 2), ("many_warehouses_2x10", 19, 7, 3)])
def test_persistent_short_horizon_against_oracle(persist, name, n, T, ignore):
    be = persist
    meta, g = G.load("rollout", name)
    data = D.slice_batch(g["data"], n)
    launches0 = be.lib.hdpo_kernel_launch_count()
    out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    # one launch per direction + set-up / weight-gradient kernels: far below the (4 layers + head) x T x 2 of the chain
    assert be.lib.hdpo_kernel_launch_count() - launches0 < 60
    fwd, flat = _oracle(meta, g, data, T)
    scale = np.abs(fwd["reward_tb"]).max()
    assert np.abs(out["reward_tb"] - fwd["reward_tb"]).max() <= 1e-5 * scale
    np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
    np.testing.assert_allclose(out["report_b"], fwd["reward_tb"][ignore:].sum(0), rtol=1e-5, atol=1e-5 * scale)
    mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    assert G.rel_l2(mine, want) <= 2e-5, G.rel_l2(mine, want)
    for k in ("store", "wh"):
        np.testing.assert_allclose(out["final"][k], fwd["final"][k], rtol=1e-4, atol=1e-4)


def test_persistent_matches_reference_golden_50_periods(persist):
    """Full 50-period golden of the unmodified reference (one_warehouse_lost_demand, 5 stores)."""
    be = persist
    meta, g = G.load("rollout", "one_warehouse_s5")
    out = D.rollout(be, meta, g["param"], g["data"], precision="tf32x3")
    T, ignore = meta["T"], meta["ignore_periods"]
    ref, ref64 = g["ref"], g["ref64"]
    true_tb = ref64["reward_tb"][:T]
    true_b = true_tb.sum(0)
    floor = np.abs(ref["reward_tb"][:T].astype(np.float64).sum(0) / true_b - 1).max()
    tol = max(1e-5, 3 * floor)  # same bar as the default path (test_kernels_abi.check_rollout_against_golden)
    assert np.abs(out["cost_b"].astype(np.float64) / true_b - 1).max() <= tol
    assert abs(out["totals"][0] - true_b.sum()) <= tol * abs(true_b.sum())
    assert np.abs(out["reward_tb"] - true_tb).max() <= 10 * tol * np.abs(true_tb).max()
    assert np.abs(out["report_b"].astype(np.float64) - true_tb[ignore:].sum(0)).max() <= tol * np.abs(true_b).max()


@pytest.mark.parametrize("layout", [K.DEMAND_BST, K.DEMAND_TSB])
def test_persistent_multi_group_batch_matches_default_path(persist, layout):
    """4096 + 300 scenarios (17 row tiles of 256 incl. a ragged one, several pair groups and head servers): the
    persistent sweeps must reproduce the default chain - same arithmetic per scenario, different GEMM tiling."""
    be = persist
    name, n, T, ignore = "one_warehouse_s5", 4096 + 300, 5, 2
    meta, g = G.load("rollout", name)
    reps = -(-n // next(iter(g["data"].values())).shape[0])
    rng = np.random.RandomState(n)
    data = {k: np.concatenate([v] * reps, 0)[:n].copy() for k, v in g["data"].items()}
    scale_b = rng.uniform(0.6, 1.4, n).astype(np.float32)
    data["demands"] *= scale_b[:, None, None]
    data["initial_inventories"] *= scale_b[:, None, None]
    new = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3", demand_layout=layout)
    be.lib.hdpo_debug_set_wide_persist(0)
    old = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3", demand_layout=layout)
    be.lib.hdpo_debug_set_wide_persist(1)
    np.testing.assert_allclose(new["cost_b"], old["cost_b"], rtol=2e-6)
    np.testing.assert_allclose(new["reward_tb"], old["reward_tb"], rtol=2e-5, atol=1e-5 * np.abs(old["reward_tb"]).max())
    assert abs(new["totals"][0] / old["totals"][0] - 1) < 1e-6
    assert G.rel_l2(new["grad_flat"], old["grad_flat"]) <= 1e-5
    for k in ("store", "wh"):
        np.testing.assert_allclose(new["final"][k], old["final"][k], rtol=1e-4, atol=1e-4)
