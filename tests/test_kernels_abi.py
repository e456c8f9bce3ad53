"""Parity of the CUDA kernels, called through the C ABI, against the reference-generated goldens and the oracle.

Each test is parametrised over two backends:
  * `cuda` (marked gpu)  - the product library on the B200: these are the parity tests proper;
  * `emu`                - the same .cu sources on the host-thread emulator, so that kernel arithmetic is
                           also checked in the CPU-only container (tests/emu; never a product path).
Tolerances: integer/index work bit-exact; per-scenario costs 1e-5 relative to the float64 reference run
(or 3x the fp32 reference's own error where that is larger); gradients relative-L2 with
the same rule.
"""
import ctypes as C

import numpy as np
import pytest

import abi_driver as D
import golden_util as G
from neural_inventory_control_b200 import _capi as K
from oracle import hdpo_oracle as O

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
_cache = {}


def backend(name):
    if name not in _cache:
        _cache[name] = D.EmuBackend() if name == "emu" else D.CudaBackend()
    return _cache[name]


SMALL = ["one_store_lost", "one_store_lost_trained", "one_store_backlogged", "one_store_backlogged_lead20",
         "serial_system", "serial_system_perturbed"]
WIDE = ["one_warehouse_s5", "one_warehouse_s50", "many_warehouses_2x10", "many_warehouses_3x50",
        # the widths bench.py measures: 153 -> 512^3 -> 51 and 309 -> 512^3 -> 153 (goldens of the unmodified reference)
        "one_warehouse_s50_w512", "many_warehouses_3x50_w512"]


def fused_cases():
    return SMALL + WIDE


def check_rollout_against_golden(out, meta, g, T, ignore):
    ref, ref64 = g["ref"], g["ref64"]
    true_tb = ref64["reward_tb"][:T]
    true_b = true_tb.sum(0)
    ref_b = ref["reward_tb"][:T].astype(np.float64).sum(0)
    floor = np.abs(ref_b / true_b - 1).max()
    tol = max(1e-5, 3 * floor)
    got_b = out["cost_b"].astype(np.float64)
    assert np.abs(got_b / true_b - 1).max() <= tol, (np.abs(got_b / true_b - 1).max(), floor)
    rep_true = true_tb[ignore:].sum(0)
    assert np.abs(out["report_b"].astype(np.float64) - rep_true).max() <= tol * np.abs(true_b).max()
    assert abs(out["totals"][0] - true_b.sum()) <= tol * abs(true_b.sum())
    assert abs(out["totals"][1] - rep_true.sum()) <= tol * abs(true_b.sum())
    # per-period per-scenario costs
    scale = np.abs(true_tb).max()
    assert np.abs(out["reward_tb"] - true_tb).max() <= 10 * tol * scale


@pytest.mark.parametrize("be_name", BACKENDS)
@pytest.mark.parametrize("name", G.step_cases())
def test_step_fwd_bwd_matches_reference(be_name, name):
    be = backend(be_name)
    meta, g = G.load("step", name)
    out = D.step(be, meta, g["data"], g["action"], up={**g["up"]}, t=meta["period"])
    ref = g["ref"]
    np.testing.assert_allclose(out["reward"], ref["reward"], rtol=2e-6, atol=1e-6)
    names = {"store": "store_inventories", "wh": "warehouse_inventories", "ech": "echelon_inventories"}
    inv = {"store": "initial_inventories", "wh": "initial_warehouse_inventories", "ech": "initial_echelon_inventories"}
    for k, v in out["new"].items():
        if v is not None:
            np.testing.assert_allclose(v, ref[f"new/{names[k]}"], rtol=2e-6, atol=2e-6)
    for k, v in out["g_cur"].items():
        if v is not None:
            np.testing.assert_allclose(v, ref[f"grad/{inv[k]}"], rtol=1e-5, atol=1e-5)
    for k, v in out["g_act"].items():
        np.testing.assert_allclose(v, ref[f"grad/action_{k}"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("be_name", BACKENDS)
def test_allocation_shift_bit_exact(be_name):
    be = backend(be_name)
    for (B, n, L) in [(1, 1, 2), (16, 7, 3), (257, 50, 6), (1000, 3, 20)]:
        h = be.zeros((B, n), np.int64)
        K.check(be.lib, be.lib.hdpo_allocation_shift(be.ptr(h), B, n, L, be.stream), "hdpo_allocation_shift")
        be.sync()
        got = be.get(h)
        assert got.dtype == np.int64 and np.array_equal(got, O.allocation_shift(B, n, L))
    meta, g = G.load("step", "many_warehouses")
    B, S, L = g["data"]["initial_inventories"].shape
    h = be.zeros((B, S), np.int64)
    K.check(be.lib, be.lib.hdpo_allocation_shift(be.ptr(h), B, S, L, be.stream), "hdpo_allocation_shift")
    be.sync()
    assert np.array_equal(be.get(h), g["ref"]["allocation_shift"])  # the reference's own table


MAPPINGS = ["unit", "unit2", "unit4", "scenario"]


class small_mapping:
    """Force one of the thread mappings of the small-net rollout: "unit" = lane = hidden unit, one scenario per warp
    (rollout_small_unit.cu; the default for batches of a few thousand scenarios), "unit2" / "unit4" = the same with 2 / 4
    scenarios per warp (weights loaded once for all of them, heads on lanes 0..G-1), "scenario" = 32 scenarios per warp,
    lane = scenario."""

    def __init__(self, be, mapping):
        self.be, self.mapping = be, mapping

    def __enter__(self):
        unit = self.mapping.startswith("unit")
        self.be.lib.hdpo_debug_set_small_unit((1 << 30) if unit else 0)
        self.be.lib.hdpo_debug_set_small_unit_group(int(self.mapping[4:] or 1) if unit else 0)

    def __exit__(self, *exc):
        self.be.lib.hdpo_debug_set_small_unit(-1)
        self.be.lib.hdpo_debug_set_small_unit_group(0)


SMALL_MODES = [pytest.param("emu", "fp32", id="emu"), pytest.param("cuda", "fp32", id="cuda", marks=pytest.mark.gpu),
               # adjoint with the HxH layers on warp-level tensor cores (mma.sync 3xTF32, mma32.cuh)
               pytest.param("cuda", "tf32x3", id="cuda-tf32x3", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("mapping", MAPPINGS)
@pytest.mark.parametrize("be_name,precision", SMALL_MODES)
@pytest.mark.parametrize("name", SMALL)
def test_rollout_costs_and_gradients_match_reference(be_name, precision, name, mapping):
    be = backend(be_name)
    meta, g = G.load("rollout", name)
    with small_mapping(be, mapping):
        out = D.rollout(be, meta, g["param"], g["data"], precision=precision)
    check_rollout_against_golden(out, meta, g, meta["T"], meta["ignore_periods"])
    ref, ref64 = g["ref"], g["ref64"]
    keys = sorted(out["grad"])
    mine = np.concatenate([out["grad"][k].ravel() for k in keys])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in keys])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in keys])
    floor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, 3 * floor), (G.rel_l2(mine, r64), floor)
    for k in keys:  # and per tensor
        assert G.rel_l2(out["grad"][k], ref64[f"grad/{k}"]) <= max(2e-5, 5 * floor), k
    names = {"store": "store_inventories", "wh": "warehouse_inventories", "ech": "echelon_inventories"}
    for k, v in out["final"].items():
        if v is not None:
            rf = ref[f"final/{names[k]}"]
            assert np.abs(v - rf).max() <= 1e-4 * max(1.0, np.abs(rf).max())


@pytest.mark.parametrize("be_name", BACKENDS)
@pytest.mark.parametrize("K_ckpt", [4, 7, 50])
@pytest.mark.parametrize("name", ["one_store_lost", "one_store_backlogged_lead20", "serial_system"])
def test_rollout_recomputation_checkpoints_change_nothing(be_name, name, K_ckpt):
    """checkpoint_interval K: the forward tapes the state of every K-th period, the adjoint re-runs the periods in
    between with the forward's own device functions, so costs AND gradients equal the K = 1 run while the state tape
    shrinks by ~K (T = 50 here: K = 7 leaves a ragged last segment, K = 50 a single checkpoint)."""
    be = backend(be_name)
    meta, g = G.load("rollout", name)
    with small_mapping(be, "scenario"):  # checkpoints exist in the 32-scenarios-per-warp form only
        full = D.rollout(be, meta, g["param"], g["data"])
        ck = D.rollout(be, meta, g["param"], g["data"], checkpoint_interval=K_ckpt)
    np.testing.assert_array_equal(ck["reward_tb"], full["reward_tb"])
    np.testing.assert_array_equal(ck["cost_b"], full["cost_b"])
    scale = np.abs(full["grad_flat"]).max()
    assert np.abs(ck["grad_flat"] - full["grad_flat"]).max() <= 1e-6 * scale


def test_checkpointed_workspace_at_bench_size():
    """Workspace of the small path at the bench size (2^20 scenarios x 50 periods), no compute: K = 1 tapes
    50 state rows per scenario, K = 10 five, plus per-warp rings that do not grow with the batch."""
    lib = backend("emu").lib
    meta, g = G.load("rollout", "one_store_backlogged_lead20")
    shapes = D.flat_params(g["param"])[1]
    net = (D.spec.mlp_widths(shapes), meta["inner_layer_activations"]["master"], meta["output_layer_activation"]["master"])
    B, T = 1 << 20, 50
    S, L = g["data"]["initial_inventories"].shape[1:]
    pb = D.spec.problem(B, S, 0, 0, L, 0, 0, meta["problem_params"]["lost_demand"], False, False)
    size = {}
    for k in (1, 10):
        desc = D.spec.rollout_desc(meta["nn_name"], pb, T, T, net, save_for_backward=True, checkpoint_interval=k,
                                   warehouse_upper_bound=meta["warehouse_upper_bound"])
        size[k] = lib.hdpo_rollout_workspace_bytes(C.byref(desc))
    row = 4 * 4 * (-(-(S * L) // 4))
    assert size[1] >= T * B * row
    assert size[10] <= 5 * B * row + (160 << 20), size  # + 4096 warp rings of 10 x 32 rows + gradient slabs
    assert size[10] < (1 << 30)


@pytest.mark.parametrize("be_name", BACKENDS)
@pytest.mark.parametrize("name", ["one_store_lost", "serial_system"])
def test_rollout_time_major_demand_layout_is_identical(be_name, name):
    be = backend(be_name)
    meta, g = G.load("rollout", name)
    a = D.rollout(be, meta, g["param"], g["data"])
    b = D.rollout(be, meta, g["param"], g["data"], demand_layout=K.DEMAND_TSB)
    assert np.array_equal(a["reward_tb"], b["reward_tb"])
    assert np.array_equal(a["grad_flat"], b["grad_flat"])


@pytest.mark.parametrize("be_name", BACKENDS)
@pytest.mark.parametrize("name,n,T,ignore", [("one_store_lost", 45, 17, 5), ("serial_system", 33, 50, 49),
                                             ("one_store_backlogged_lead20", 1, 3, 0)])
@pytest.mark.parametrize("mapping", MAPPINGS)
def test_rollout_ragged_batches_against_oracle(be_name, name, n, T, ignore, mapping, request):
    """B not a multiple of the warp tile, short horizons, ignore near T: compare with the pinned oracle."""
    be = backend(be_name)
    ctx = small_mapping(be, mapping)
    ctx.__enter__()
    request.addfinalizer(lambda: ctx.__exit__())
    meta, g = G.load("rollout", name)
    data = D.slice_batch(g["data"], n)
    out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, g_total=0.37, g_report=-0.11)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    d64 = G.cast(data, np.float64)
    fwd = O.rollout_forward(pol, pb, d64, T, ignore)
    np.testing.assert_allclose(out["reward_tb"], fwd["reward_tb"], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(out["report_b"], fwd["reward_tb"][ignore:].sum(0), rtol=2e-5, atol=2e-4)
    # gradient with two different upstream scalars == linear combination of two oracle adjoints
    _, g_all = O.rollout_grad(pol, pb, d64, T, grad_scale=1.0)
    full = O.flatten_grads(pol, g_all)
    # report-only part: rerun the oracle adjoint on the tail by zeroing r_bar before `ignore` is not available, so
    # use linearity: grad(g_total, g_report) = g_total * grad_all + g_report * grad_tail, with grad_tail from the
    # kernel itself at (0, 1), and check the (1, 0) leg against the oracle.
    leg_all = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, g_total=1.0, g_report=0.0)
    leg_tail = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, g_total=0.0, g_report=1.0)
    for k in full:
        assert G.rel_l2(leg_all["grad"][k], full[k]) < 2e-5, k
    combo = 0.37 * leg_all["grad_flat"].astype(np.float64) - 0.11 * leg_tail["grad_flat"].astype(np.float64)
    assert G.rel_l2(out["grad_flat"], combo) < 1e-5
    if ignore >= T:
        assert np.all(leg_tail["grad_flat"] == 0)


@pytest.mark.parametrize("be_name,n", [pytest.param("emu", 300, id="emu-300"),
                                       pytest.param("cuda", 300, id="cuda-300", marks=pytest.mark.gpu),
                                       pytest.param("cuda", 100000, id="cuda-100000", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("name", ["one_store_backlogged", "serial_system"])
@pytest.mark.parametrize("mapping", MAPPINGS)
def test_rollout_many_tiles_all_launch_shapes(be_name, n, name, mapping, request):
    """Batches spanning many warp tiles (1, 2 and 4 warps per CTA, persistent tile loops, ragged last tile):
    per-scenario costs of replicated scenarios must repeat exactly and the gradient must equal the oracle's."""
    be = backend(be_name)
    ctx = small_mapping(be, mapping)
    ctx.__enter__()
    request.addfinalizer(lambda: ctx.__exit__())
    meta, g = G.load("rollout", name)
    base = 50
    reps = -(-n // base)
    data = {k: np.concatenate([v[:base]] * reps, 0)[:n] for k, v in g["data"].items()}
    T, ignore = 6, 2
    out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore)
    first = out["cost_b"][:base]
    for r in range(1, n // base):
        assert np.array_equal(out["cost_b"][r * base:(r + 1) * base], first)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    small = {k: v[:base].astype(np.float64) for k, v in g["data"].items()}
    fwd, grads = O.rollout_grad(pol, pb, small, T, grad_scale=1.0)
    np.testing.assert_allclose(first, fwd["reward_tb"].sum(0), rtol=1e-5)
    # gradient of the replicated batch = (n / base) x the base gradient (up to the ragged tail), scaled by 1/(n T S)
    flat = O.flatten_grads(pol, grads)
    if n % base == 0:
        want = np.concatenate([flat[k].ravel() for k in sorted(flat)]) * (n // base) / (n * T * pb.n_stores)
        got = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
        assert G.rel_l2(got, want) < 1e-5
    assert abs(out["totals"][0] - out["cost_b"].astype(np.float64).sum()) <= 1e-9 * abs(out["totals"][0])


@pytest.mark.parametrize("be_name", BACKENDS)
def test_discrete_allocation_forward(be_name):
    be = backend(be_name)
    meta, g = G.load("rollout", "one_store_lost")
    out = D.rollout(be, meta, g["param"], g["data"], T=20, discrete=True, backward=False)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float32)
    fwd = O.rollout_forward(pol, pb, G.cast(g["data"], np.float32), 20, meta["ignore_periods"], discrete=True)
    # rounding makes the trajectory piecewise constant in the weights: identical unless an action sits on a .5 tie
    close = np.isclose(out["reward_tb"], fwd["reward_tb"], rtol=1e-5, atol=1e-4)
    assert close.mean() > 0.99


@pytest.mark.parametrize("be_name", BACKENDS)
def test_backward_rejects_discrete_and_small_workspace(be_name):
    be = backend(be_name)
    meta, g = G.load("rollout", "one_store_lost")
    with pytest.raises(K.HdpoError):
        D.rollout(be, meta, g["param"], g["data"], T=5, discrete=True, backward=True)


# ---- wide path (VanillaWarehouse: tile-GEMM pipeline + warehouse head kernels) ----
def _grad_check_vs_golden(out, g, floor_mult=3):
    ref, ref64 = g["ref"], g["ref64"]
    keys = sorted(out["grad"])
    mine = np.concatenate([out["grad"][k].ravel() for k in keys])
    r32 = np.concatenate([ref[f"grad/{k}"].ravel() for k in keys])
    r64 = np.concatenate([ref64[f"grad/{k}"].ravel() for k in keys])
    floor = G.rel_l2(r32, r64)
    assert G.rel_l2(mine, r64) <= max(1e-5, floor_mult * floor), (G.rel_l2(mine, r64), floor)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", WIDE)
def test_wide_rollout_costs_and_gradients_match_reference(name, precision):
    """Full 50-period goldens of the warehouse settings (GPU only: the emulator needs ~30 s per case), through the
    fp32 SIMT GEMM and through the tcgen05 3xTF32 GEMM - both are parity modes and must meet the same bar."""
    be = backend("cuda")
    meta, g = G.load("rollout", name)
    out = D.rollout(be, meta, g["param"], g["data"], precision=precision)
    check_rollout_against_golden(out, meta, g, meta["T"], meta["ignore_periods"])
    # gradient against the float64 run of the reference: 1e-5, or 3x the fp32 reference's own distance to it where that
    # is larger (50-period warehouse rollouts are chaotic) - the same bar for the fp32 and the 3xTF32 mode
    _grad_check_vs_golden(out, g, floor_mult=3)
    names = {"store": "store_inventories", "wh": "warehouse_inventories"}
    for k, rk in names.items():
        rf = g["ref"][f"final/{rk}"]
        assert np.abs(out["final"][k] - rf).max() <= 2e-3 * max(1.0, np.abs(rf).max())


WIDE_MODES = [pytest.param("emu", "fp32", id="emu-fp32"),
              pytest.param("cuda", "fp32", id="cuda-fp32", marks=pytest.mark.gpu),
              pytest.param("cuda", "tf32x3", id="cuda-tf32x3", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("be_name,precision", WIDE_MODES)
@pytest.mark.parametrize("name,n,T,ignore", [("one_warehouse_s5", 32, 6, 2), ("many_warehouses_2x10", 19, 7, 3),
                                             ("many_warehouses_3x50", 5, 4, 0)])
def test_wide_rollout_short_horizon_against_oracle(be_name, precision, name, n, T, ignore):
    """Short horizons (before the chaotic regime) against the pinned float64 oracle: tight tolerances."""
    be = backend(be_name)
    meta, g = G.load("rollout", name)
    data = D.slice_batch(g["data"], n)
    out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision=precision)
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    fwd, grads = O.rollout_grad(pol, pb, G.cast(data, np.float64), T)
    scale = np.abs(fwd["reward_tb"]).max()
    assert np.abs(out["reward_tb"] - fwd["reward_tb"]).max() <= 1e-5 * scale
    np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
    np.testing.assert_allclose(out["report_b"], fwd["reward_tb"][ignore:].sum(0), rtol=1e-5, atol=1e-5 * scale)
    flat = O.flatten_grads(pol, grads)
    mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    assert G.rel_l2(mine, want) <= 2e-5, G.rel_l2(mine, want)
    for k in ("store", "wh"):
        np.testing.assert_allclose(out["final"][k], fwd["final"][k], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_wide_rollout_single_pass_tf32_is_close_but_not_parity_grade():
    """HDPO_PREC_TF32 (throughput mode) runs the same pipeline with one tensor pass: 10-bit-mantissa accuracy."""
    be = backend("cuda")
    meta, g = G.load("rollout", "many_warehouses_2x10")
    data = D.slice_batch(g["data"], 32)
    fast = D.rollout(be, meta, g["param"], data, T=8, ignore=0, precision="tf32")
    exact = D.rollout(be, meta, g["param"], data, T=8, ignore=0, precision="fp32")
    rel = np.abs(fast["cost_b"] / exact["cost_b"] - 1).max()
    assert 1e-7 < rel < 2e-2, rel
    assert G.rel_l2(fast["grad_flat"], exact["grad_flat"]) < 5e-2


@pytest.mark.gpu
def test_wide_rollout_512_wide_tc_vs_simt_and_padding_rows():
    """B not a multiple of the 128-row tile and a NaN-poisoned workspace: tile-padding rows must not leak into the
    weight gradient. 512-wide net (K = 512 accumulations per output): both parity modes are compared with the
    float64 oracle; the tcgen05 3xTF32 error must be fp32-grade, i.e. within 3x of what the fp32 SIMT path achieves
    (the short-horizon gradient is ill-conditioned, so the yardstick is the fp32 path's own error, not 1e-5)."""
    be = backend("cuda")
    meta, g = G.load("rollout", "one_warehouse_s50")
    rng = np.random.RandomState(0)
    params = {}
    widths = [153, 512, 512, 51]
    for i in range(3):
        k = 1 / np.sqrt(widths[i])
        params[f"net.master.{2 * i}.weight"] = rng.uniform(-k, k, (widths[i + 1], widths[i])).astype(np.float32)
        params[f"net.master.{2 * i}.bias"] = rng.uniform(-k, k, (widths[i + 1],)).astype(np.float32)
    data = {k: np.concatenate([v] * 13, 0)[:200] for k, v in g["data"].items()}  # 200 scenarios -> 56 padding rows
    T = 12
    meta = dict(meta, neurons_per_hidden_layer={"master": widths[1:-1]})
    pol = G.policy_from_golden(meta, params, np.float64)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(data, np.float64), T)
    flat = O.flatten_grads(pol, grads)
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    errs = {}
    for prec in ("fp32", "tf32x3"):
        out = D.rollout(be, meta, params, data, T=T, ignore=1, precision=prec)
        assert np.isfinite(out["grad_flat"]).all()
        np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
        got = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
        errs[prec] = G.rel_l2(got, want)
    print("512-wide gradient rel-L2 error vs float64:", errs)
    assert errs["tf32x3"] <= max(1e-5, 3 * errs["fp32"]), errs


@pytest.mark.gpu
@pytest.mark.parametrize("name,precision", [("one_store_lost", "fp32"), ("serial_system", "fp32"),
                                            ("many_warehouses_2x10", "tf32x3")])
def test_rollout_train_host_entry_point(name, precision):
    """hdpo_rollout_train_host (HOST buffers in, totals + gradient out; what bench.py's `e2e` times) returns the
    same numbers as the device-pointer forward + adjoint pair with dLoss/dtotal = 1/(B*T*S)."""
    import ctypes as C
    import torch
    be = backend("cuda")
    meta, g = G.load("rollout", name)
    ref = D.rollout(be, meta, g["param"], g["data"], precision=precision)
    pp = meta["problem_params"]
    host = {k: np.ascontiguousarray(v, np.float32) for k, v in g["data"].items()}
    flat, shapes, names = D.flat_params(g["param"])
    widths = D.spec.mlp_widths(shapes)
    B, S, L = host["initial_inventories"].shape
    W, E = pp["n_warehouses"], pp["n_extra_echelons"]
    Lw = host["initial_warehouse_inventories"].shape[2] if W else 0
    Le = host["initial_echelon_inventories"].shape[2] if E else 0
    pb = D.spec.problem(B, S, W, E, L, Lw, Le, pp["lost_demand"], pp["maximize_profit"],
                        host.get("warehouse_edge_costs") is not None)
    adj = pp.get("warehouse_store_adjacency")
    h_adj = None if adj is None else np.ascontiguousarray(np.asarray(adj), np.int32)
    desc = D.spec.rollout_desc(meta["nn_name"], pb, meta["T"], host["demands"].shape[2],
                               (widths, meta["inner_layer_activations"]["master"], None),
                               ignore_periods=meta["ignore_periods"], precision=precision,
                               warehouse_upper_bound=meta["warehouse_upper_bound"])
    hp = lambda k: host[k].ctypes.data if k in host else None  # noqa: E731
    st = K.Statics(hp("holding_costs"), hp("underage_costs"), hp("lead_times"), hp("warehouse_lead_times"),
                   hp("warehouse_holding_costs"), hp("warehouse_edge_costs"), hp("echelon_lead_times"),
                   hp("echelon_holding_costs"), hp("mean"), hp("std"))
    init = K.State(hp("initial_inventories"), hp("initial_warehouse_inventories"), hp("initial_echelon_inventories"))
    ws_bytes = be.lib.hdpo_rollout_host_workspace_bytes(C.byref(desc))
    assert ws_bytes > 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    totals = np.zeros(2, np.float64)
    grad = np.full(flat.size, np.nan, np.float32)
    rc = be.lib.hdpo_rollout_train_host(C.byref(desc), flat.ctypes.data, host["demands"].ctypes.data, C.byref(st),
                                        C.byref(init), None if h_adj is None else h_adj.ctypes.data,
                                        totals.ctypes.data, grad.ctypes.data, ws.data_ptr(), ws_bytes, None)
    K.check(be.lib, rc, "hdpo_rollout_train_host")
    np.testing.assert_allclose(totals, ref["totals"], rtol=1e-6)
    assert G.rel_l2(grad, ref["grad_flat"]) < 1e-6
    # too-small workspace is refused, not overrun
    rc = be.lib.hdpo_rollout_train_host(C.byref(desc), flat.ctypes.data, host["demands"].ctypes.data, C.byref(st),
                                        C.byref(init), None if h_adj is None else h_adj.ctypes.data,
                                        totals.ctypes.data, grad.ctypes.data, ws.data_ptr(), 1024, None)
    assert rc == -4


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name,n,T,ignore", [("many_warehouses_2x10", 4096 + 300, 5, 2), ("one_warehouse_s5", 4096, 4, 0)])
def test_wide_rollout_concurrent_chunks_against_oracle(precision, name, n, T, ignore):
    """Batches of >= 4096 scenarios are cut into chunks that run on concurrent streams (rollout_wide.cu "host
    orchestration"); the ragged case leaves the second chunk with a partial row tile. Costs, per-period rewards,
    final state and the fixed-order-summed gradient must match the float64 oracle exactly as for one chunk."""
    be = backend("cuda")
    meta, g = G.load("rollout", name)
    reps = -(-n // next(iter(g["data"].values())).shape[0])
    rng = np.random.RandomState(n)
    data = {k: np.concatenate([v] * reps, 0)[:n].copy() for k, v in g["data"].items()}
    # de-duplicate the tiled scenarios: scale demands and initial inventories per scenario
    scale_b = rng.uniform(0.6, 1.4, n).astype(np.float32)
    data["demands"] *= scale_b[:, None, None]
    data["initial_inventories"] *= scale_b[:, None, None]
    pb = G.problem_from_meta(meta)
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    fwd, grads = O.rollout_grad(pol, pb, G.cast(data, np.float64), T)
    flat = O.flatten_grads(pol, grads)
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    scale = np.abs(fwd["reward_tb"]).max()
    for layout in (K.DEMAND_BST, K.DEMAND_TSB):
        out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision=precision, demand_layout=layout)
        assert np.abs(out["reward_tb"] - fwd["reward_tb"]).max() <= 1e-5 * scale, layout
        np.testing.assert_allclose(out["cost_b"], fwd["reward_tb"].sum(0), rtol=1e-5)
        np.testing.assert_allclose(out["report_b"], fwd["reward_tb"][ignore:].sum(0), rtol=1e-5, atol=1e-5 * scale)
        assert abs(out["totals"][0] / fwd["reward_tb"].sum() - 1) < 1e-6
        mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
        assert G.rel_l2(mine, want) <= 2e-5, (layout, G.rel_l2(mine, want))
        for k in ("store", "wh"):
            np.testing.assert_allclose(out["final"][k], fwd["final"][k], rtol=1e-4, atol=1e-4)
