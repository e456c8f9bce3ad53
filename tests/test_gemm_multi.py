"""Multi-tile CTA-pair GEMM of the wide path (csrc/wide_persist.cu, `wp::gemm_multi`) through the C ABI. GPU only.

Layer launches with enough 256-row tiles are routed to a kernel in which every CTA pair walks several tiles, the
epilogue of one overlapping the MMAs of the next. `hdpo_debug_set_tc_multi(1)` routes EVERY tensor-core GEMM there so
that the small goldens exercise all four epilogues (hidden / output forward, hidden / state-adjoint dgrad):
  * the raw product against float64 numpy (3xTF32 fp32-grade, single-pass TF32 at its own accuracy);
  * the reference's goldens at the bench widths and a short-horizon float64 oracle run, same bars as the default path;
  * a chunked 4396-scenario batch against the single-tile forms.
"""
import numpy as np
import pytest

import abi_driver as D
import golden_util as G
from neural_inventory_control_b200 import _capi as K
from oracle import hdpo_oracle as O
from test_kernels_abi import check_rollout_against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def multi():
    be = D.CudaBackend()
    be.lib.hdpo_debug_set_tc_multi(1)
    try:
        yield be
    finally:
        be.lib.hdpo_debug_set_tc_multi(-1)


# 256 x 64 | 128 tiles; (4096, 512): 64 tiles = one per pair; (8192, 512): two per pair; (19 * 256, 192): ragged walk
@pytest.mark.parametrize("M,N,Kd", [(256, 64, 32), (256, 128, 64), (512, 192, 96), (4096, 512, 512), (8192, 512, 512),
                                    (19 * 256, 192, 512), (8192, 64, 512), (2048, 512, 192), (16384, 256, 64)])
def test_gemm_multi_matches_float64(multi, M, N, Kd):
    be = multi
    rng = np.random.RandomState(M + N + Kd)
    A = rng.randn(M, Kd).astype(np.float32)
    B = (rng.randn(N, Kd) / np.sqrt(Kd)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(want).max()
    a, b = be.put(A), be.put(B)
    scratch = be.zeros(2 * (M * Kd + N * Kd))
    fp32_err = np.abs((A @ B.T).astype(np.float64) - want).max() / scale
    for n_pass, tol in ((3, max(2e-6, 4 * fp32_err)), (1, 2e-3)):
        c = be.put(np.full((M, N), np.nan, np.float32))
        rc = be.lib.hdpo_debug_gemm_tc(be.ptr(a), be.ptr(b), be.ptr(c), M, N, Kd, n_pass, be.ptr(scratch), be.stream)
        K.check(be.lib, rc, "hdpo_debug_gemm_tc")
        be.sync()
        err = np.abs(be.get(c).astype(np.float64) - want).max() / scale
        print(f"gemm_multi {M}x{N}x{Kd} n_pass={n_pass}: max err/scale {err:.3e} (fp32 numpy: {fp32_err:.3e})")
        assert err < tol, (n_pass, err)
        if n_pass == 1:
            assert err > 1e-6


@pytest.mark.parametrize("precision", ["tf32x3"])
@pytest.mark.parametrize("name", ["one_warehouse_s5", "one_warehouse_s50_w512", "many_warehouses_3x50_w512"])
def test_multi_rollout_matches_reference_golden(multi, name, precision):
    be = multi
    meta, g = G.load("rollout", name)
    T, ignore = meta["T"], meta["ignore_periods"]
    out = D.rollout(be, meta, g["param"], g["data"], precision=precision)
    check_rollout_against_golden(out, meta, g, T, ignore)
    ref, ref64 = g["ref"], g["ref64"]
    keys = sorted(out["grad"])
    true = np.concatenate([np.asarray(ref64[f"grad/{k}"], np.float64).ravel() for k in keys])
    theirs = np.concatenate([np.asarray(ref[f"grad/{k}"], np.float64).ravel() for k in keys])
    mine = np.concatenate([out["grad"][k].ravel() for k in keys])
    floor = G.rel_l2(theirs, true)
    assert G.rel_l2(mine, true) <= max(1e-5, 3 * floor), (G.rel_l2(mine, true), floor)


@pytest.mark.parametrize("name,n,T,ignore", [("one_warehouse_s5", 300, 6, 2), ("many_warehouses_2x10", 19, 7, 3)])
def test_multi_rollout_short_horizon_against_oracle(multi, name, n, T, ignore):
    be = multi
    meta, g = G.load("rollout", name)
    reps = -(-n // next(iter(g["data"].values())).shape[0])
    data = {k: np.concatenate([v] * reps, 0)[:n].copy() for k, v in g["data"].items()}
    rng = np.random.RandomState(n)
    s = rng.uniform(0.6, 1.4, n).astype(np.float32)
    data["demands"] *= s[:, None, None]
    out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    fwd, grads = O.rollout_grad(pol, pb, G.cast(data, np.float64), T)
    flat = O.flatten_grads(pol, grads)
    scale = np.abs(fwd["reward_tb"]).max()
    assert np.abs(out["reward_tb"] - fwd["reward_tb"]).max() <= 1e-5 * scale
    np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
    mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    assert G.rel_l2(mine, want) <= 2e-5, G.rel_l2(mine, want)


def test_multi_chunked_batch_matches_single_tile_forms(multi):
    """4096 + 300 scenarios: chunks on concurrent streams, a ragged last row tile; same arithmetic per scenario as the
    single-tile GEMM forms, different tiling and accumulation segments."""
    be = multi
    name, n, T, ignore = "one_warehouse_s5", 4096 + 300, 5, 2
    meta, g = G.load("rollout", name)
    reps = -(-n // next(iter(g["data"].values())).shape[0])
    rng = np.random.RandomState(n)
    data = {k: np.concatenate([v] * reps, 0)[:n].copy() for k, v in g["data"].items()}
    scale_b = rng.uniform(0.6, 1.4, n).astype(np.float32)
    data["demands"] *= scale_b[:, None, None]
    data["initial_inventories"] *= scale_b[:, None, None]
    l0 = be.lib.hdpo_kernel_launch_count()
    new = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    l1 = be.lib.hdpo_kernel_launch_count()
    be.lib.hdpo_debug_set_tc_multi(0)
    old = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    be.lib.hdpo_debug_set_tc_multi(1)
    assert l1 > l0
    np.testing.assert_allclose(new["cost_b"], old["cost_b"], rtol=2e-6)
    np.testing.assert_allclose(new["reward_tb"], old["reward_tb"], rtol=2e-5, atol=1e-5 * np.abs(old["reward_tb"]).max())
    assert abs(new["totals"][0] / old["totals"][0] - 1) < 1e-6
    assert G.rel_l2(new["grad_flat"], old["grad_flat"]) <= 1e-5
    for k in ("store", "wh"):
        np.testing.assert_allclose(new["final"][k], old["final"][k], rtol=1e-4, atol=1e-4)
