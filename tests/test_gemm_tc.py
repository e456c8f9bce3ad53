"""tcgen05 / TMEM / TMA tile GEMM (csrc/gemm_tc.cu) against a float64 numpy product. GPU only (tensor cores cannot
be emulated). 3xTF32 must be fp32-grade (1e-5 bar of the parity mode); single-pass TF32 is checked at its own
10-bit-mantissa accuracy so that a silently wrong descriptor / swizzle cannot hide behind a loose tolerance."""
import numpy as np
import pytest

import abi_driver as D
from neural_inventory_control_b200 import _capi as K

pytestmark = pytest.mark.gpu


# M % 256 == 0 and N % 128 == 0 select the CTA-pair (cta_group::2) form, everything else the single-CTA tiles
@pytest.mark.parametrize("M,N,Kd", [(128, 64, 32), (128, 128, 64), (256, 192, 96), (384, 512, 512), (1024, 64, 512),
                                    (2048, 512, 192), (256, 128, 32), (512, 256, 512), (2048, 512, 512)])
def test_gemm_tc_matches_float64(M, N, Kd):
    be = D.CudaBackend()
    rng = np.random.RandomState(M + N + Kd)
    A = rng.randn(M, Kd).astype(np.float32)
    B = (rng.randn(N, Kd) / np.sqrt(Kd)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(want).max()
    a, b = be.put(A), be.put(B)
    scratch = be.zeros(2 * (M * Kd + N * Kd))
    fp32_err = np.abs((A @ B.T).astype(np.float64) - want).max() / scale  # what a true-fp32 GEMM achieves here
    for n_pass, tol in ((3, max(2e-6, 4 * fp32_err)), (1, 2e-3)):
        c = be.zeros((M, N))
        rc = be.lib.hdpo_debug_gemm_tc(be.ptr(a), be.ptr(b), be.ptr(c), M, N, Kd, n_pass, be.ptr(scratch), be.stream)
        K.check(be.lib, rc, "hdpo_debug_gemm_tc")
        be.sync()
        got = be.get(c).astype(np.float64)
        err = np.abs(got - want).max() / scale
        print(f"gemm_tc {M}x{N}x{Kd} n_pass={n_pass}: max err/scale {err:.3e} (fp32 numpy: {fp32_err:.3e})")
        assert err < tol, (n_pass, err)
        if n_pass == 1:
            assert err > 1e-6  # it really is the single-pass path


def test_gemm_tc_rejects_untileable_shapes():
    be = D.CudaBackend()
    z = be.zeros(16)
    assert be.lib.hdpo_debug_gemm_tc(be.ptr(z), be.ptr(z), be.ptr(z), 100, 64, 32, 3, be.ptr(z), be.stream) != 0
    assert b"tileable" in be.lib.hdpo_last_error()


@pytest.mark.parametrize("M,N,Kd,kps", [(128, 64, 64, 32), (128, 128, 256, 128), (256, 192, 1024, 256),
                                        (512, 512, 4096, 1024), (64 * 2, 64, 2048, 1024)])
def test_gemm_tc_weight_gradient_form(M, N, Kd, kps):
    """MN-major operands straight from [row][feature] tapes: C = A^T B with split-K partial slices."""
    be = D.CudaBackend()
    rng = np.random.RandomState(M + N + Kd)
    A = rng.randn(Kd, M).astype(np.float32)
    B = (rng.randn(Kd, N) / np.sqrt(Kd)).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    scale = np.abs(want).max()
    fp32_err = np.abs((A.T @ B).astype(np.float64) - want).max() / scale
    a, b = be.put(A), be.put(B)
    scratch = be.zeros(2 * (Kd * M + Kd * N) + (Kd // kps) * M * N)
    for n_pass, tol in ((3, max(2e-6, 4 * fp32_err)), (1, 2e-3)):
        c = be.zeros((M, N))
        rc = be.lib.hdpo_debug_gemm_tc_wgrad(be.ptr(a), be.ptr(b), be.ptr(c), M, N, Kd, kps, n_pass, be.ptr(scratch),
                                             be.stream)
        K.check(be.lib, rc, "hdpo_debug_gemm_tc_wgrad")
        be.sync()
        err = np.abs(be.get(c).astype(np.float64) - want).max() / scale
        print(f"gemm_tc wgrad {M}x{N}x{Kd}/{kps} n_pass={n_pass}: max err/scale {err:.3e} (fp32 numpy: {fp32_err:.3e})")
        assert err < tol, (n_pass, err)


@pytest.mark.parametrize("M,N,Kd", [(2048, 512, 512), (256, 128, 64), (384, 64, 96)])
def test_gemm_tc_forward_hidden_epilogue_hook(M, N, Kd):
    """hdpo_debug_gemm_tc_timeline with epi = 0 is the kernel bench.py times for `roofline`: the forward hidden-layer
    form, (hi, lo) = split(ELU(A B^T + bias)). hi + lo must be the fp32-grade result and hi its tf32 rounding."""
    be = D.CudaBackend()
    rng = np.random.RandomState(7 + M + N + Kd)
    A = rng.randn(M, Kd).astype(np.float32)
    B = (rng.randn(N, Kd) / np.sqrt(Kd)).astype(np.float32)
    bias = rng.randn(N).astype(np.float32)
    z = A.astype(np.float64) @ B.astype(np.float64).T + bias.astype(np.float64)
    want = np.where(z > 0, z, np.expm1(z))
    a, b, bs = be.put(A), be.put(B), be.put(bias)
    scratch = be.zeros(2 * (M * Kd + N * Kd))
    c, c_lo = be.zeros((M, N)), be.zeros((M, N))
    rc = be.lib.hdpo_debug_gemm_tc(be.ptr(a), be.ptr(b), be.ptr(c), M, N, Kd, 3, be.ptr(scratch), be.stream)  # splits
    K.check(be.lib, rc, "hdpo_debug_gemm_tc")
    dbg = be.zeros(8 * 4096, np.int64)
    rc = be.lib.hdpo_debug_gemm_tc_timeline(be.ptr(a), be.ptr(b), be.ptr(c), M, N, Kd, 3, be.ptr(scratch), be.ptr(dbg),
                                            be.stream, 0, be.ptr(c_lo), be.ptr(bs))
    K.check(be.lib, rc, "hdpo_debug_gemm_tc_timeline")
    be.sync()
    hi, lo = be.get(c).astype(np.float64), be.get(c_lo).astype(np.float64)
    err = np.abs(hi + lo - want).max() / np.abs(want).max()
    assert err < 3e-6, err
    assert np.abs(lo).max() <= 2.0 ** -10 * np.abs(hi).max()          # lo is the remainder of a tf32 rounding
    assert np.all((be.get(c).view(np.uint32) & 0x1FFF) == 0)          # hi carries 10 mantissa bits
    assert np.count_nonzero(be.get(dbg)) > 0                          # the per-CTA clock stamps were written
