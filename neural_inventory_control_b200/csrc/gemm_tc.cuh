// gemm_tc.cuh - tcgen05 (5th-gen tensor core) tile GEMM for the wide policy MLP, sm_100a only.
//
//   D[M,N] = epi( A[M,K] * B[N,K]^T )     both operands K-major fp32 rows in global memory
//
// * operands are moved by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 3/4-stage shared-memory ring,
//   full/empty mbarriers, one producer lane; one MMA lane issues tcgen05.mma.cta_group::1.kind::tf32
//   (M = 128, N = BN, K = 8 per instruction) into a TMEM accumulator (BN fp32 columns x 128 lanes);
//   tcgen05.commit releases ring slots and finally signals the four epilogue warps, which read the accumulator
//   with tcgen05.ld (32x32b: lane = output row) and apply the fused epilogue.
// * precision: "3xTF32" - every fp32 operand x is pre-split by the producing kernel into hi = tf32(x) and
//   lo = x - hi (exact); per K step the MMA lane issues hi*hi + lo*hi + hi*lo into the same fp32 accumulator,
//   which restores ~fp32 accuracy (error ~2^-21) at 3x the tensor time. HDPO_PREC_TF32 issues hi*hi only.
//   Epilogues therefore WRITE hi/lo pairs for everything a later GEMM consumes.
#pragma once

#include "hdpo_internal.cuh"

#ifndef HDPO_EMU
#include <cuda.h>

namespace hdpo {
namespace tc {

enum { EPI_FWD_HIDDEN = 0, EPI_FWD_OUT = 1, EPI_DGRAD_HIDDEN = 2, EPI_DGRAD_ACCUM = 3, EPI_STORE = 4 };

struct GemmTcArgs {
  int M, N, K;          // M % 128 == 0, N % BN == 0, K % 32 == 0
  int k_per_split;      // weight-gradient form only: contraction rows handled by one blockIdx.z slice
  size_t c_slice;       // weight-gradient form only: floats between the partial outputs of consecutive slices
  int n_pass;           // 3 = 3xTF32, 1 = single-pass TF32
  int hi_chunks;        // accumulators the hi*hi term is spread over (0 = default, see gemm_tc.cu)
  int a_row0, b_row0;   // row origin of this GEMM inside the A / B tensor maps (e.g. t * Bp for tapes)
  int ldc;              // leading dimension (floats) of every output / aux array
  int act;              // HDPO_ACT_* of the epilogue
  float* c_full;        // EPI_FWD_OUT / EPI_DGRAD_ACCUM / EPI_STORE
  float* c_hi;          // EPI_FWD_HIDDEN / EPI_DGRAD_HIDDEN
  float* c_lo;
  const float* bias;    // EPI_FWD_*
  const float* aux_hi;  // EPI_DGRAD_HIDDEN: saved layer output h = aux_hi + aux_lo
  const float* aux_lo;
  float* colsum_part;   // EPI_DGRAD_HIDDEN, optional: [rows / 32][ldc] column sums of each 32-row block of the output
                        // (bias-gradient partials; the row index includes a_row0, i.e. it follows the gz tape)
};

// one 2-D fp32 tensor map over a row-major [rows][ld] array, box = [box_rows][32 floats]; 128B swizzle for K-major
// operands, 128B swizzle with 32-byte atoms for the MN-major (weight-gradient) operands
int make_tensor_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    bool mn_major = false);

// weight-gradient form: D[M,N] (per K slice) = sum_k A[k][m] * B[k][n] with A = [K][M], B = [K][N] row-major arrays
// (tensor maps with box_rows = 32); partial slice z is written at c_full + z * c_slice.
int gemm_wgrad(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
               const GemmTcArgs& g, int bn, void* stream);

int gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
         const GemmTcArgs& g, int epi, int bn, void* stream);

constexpr int kBoxRowsA = 128;
#ifdef HDPO_TC_BN64_EXPERIMENT
inline int pick_bn(int) { return 64; }
#else
inline int pick_bn(int N) { return (N % 128 == 0) ? 128 : 64; }
#endif

}  // namespace tc
}  // namespace hdpo
#endif
