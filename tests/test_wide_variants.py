"""Parity of the OPT-IN launch variants of the wide path's per-period chain (rollout_wide.cu / gemm_tc.cu) through the C ABI:

  * two CTAs per SM: 256 x 64 CTA-pair tiles / 128 x 64 single-CTA tiles (`hdpo_debug_set_tc_occ2`, HDPO_TC_OCC2),
  * weight-gradient GEMMs overlapped with the adjoint sweep in period groups on a second stream, forced on / off
    (`hdpo_debug_set_wide_wg_overlap`, HDPO_WIDE_WG_OVERLAP; the default turns it on for one-chunk batches).

Every variant computes the same tiles from the same operands, so it must reproduce the default path to rounding level and the
pinned float64 oracle at the bars of test_kernels_abi.py. The batch is the golden's 32 scenarios repeated 8 times (256 rows:
the CTA-pair forms need whole 256-row tiles); with dLoss/dtotal = 1/(B T S) its gradient equals the 32-scenario gradient.
"""
import numpy as np
import pytest

import abi_driver as D
import golden_util as G
from oracle import hdpo_oracle as O

pytestmark = pytest.mark.gpu

VARIANTS = {
    "occ2_pairs": lambda lib: lib.hdpo_debug_set_tc_occ2(1),
    "occ2_single": lambda lib: lib.hdpo_debug_set_tc_occ2(2),
    "wg_overlap_groups_of_2": lambda lib: lib.hdpo_debug_set_wide_wg_overlap(1, 2),
    "wg_overlap_off": lambda lib: lib.hdpo_debug_set_wide_wg_overlap(0, 0),
}


def _reset(lib):
    lib.hdpo_debug_set_tc_occ2(0)
    lib.hdpo_debug_set_wide_wg_overlap(-1, 5)


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("name,T,ignore", [("one_warehouse_s5", 6, 2), ("many_warehouses_2x10", 7, 3)])
def test_variant_matches_default_path_and_oracle(variant, name, T, ignore):
    be = D.CudaBackend()
    meta, g = G.load("rollout", name)
    small = D.slice_batch(g["data"], 32)
    data = {k: np.concatenate([v] * 8, axis=0) for k, v in small.items()}
    _reset(be.lib)
    base = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    try:
        VARIANTS[variant](be.lib)
        out = D.rollout(be, meta, g["param"], data, T=T, ignore=ignore, precision="tf32x3")
    finally:
        _reset(be.lib)
    scale = np.abs(base["reward_tb"]).max()
    assert np.abs(out["reward_tb"] - base["reward_tb"]).max() <= 1e-6 * scale
    assert G.rel_l2(out["grad_flat"], base["grad_flat"]) <= 1e-6
    # and the oracle on the 32 distinct scenarios
    pol = G.policy_from_golden(meta, g["param"], np.float64)
    fwd, grads = O.rollout_grad(pol, G.problem_from_meta(meta), G.cast(small, np.float64), T)
    want_tb = np.concatenate([fwd["reward_tb"]] * 8, axis=1)
    assert np.abs(out["reward_tb"] - want_tb).max() <= 1e-5 * np.abs(want_tb).max()
    flat = O.flatten_grads(pol, grads)
    mine = np.concatenate([out["grad"][k].ravel() for k in sorted(flat)])
    want = np.concatenate([flat[k].ravel() for k in sorted(flat)])
    assert G.rel_l2(mine, want) <= 2e-5, G.rel_l2(mine, want)
