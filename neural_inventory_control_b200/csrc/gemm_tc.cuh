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
  int k_per_split;      // weight-gradient form: contraction rows handled by one blockIdx.z slice. K-major single-CTA
                        // forms: > 0 = split-K, slice z takes columns [z, z + 1) * k_per_split of A / B and writes its
                        // partial product c_zrows rows further down the c0 map (bias added by slice 0 only)
  int c_zrows;          // K-major split-K: rows between the partial outputs of consecutive slices
  size_t c_slice;       // weight-gradient form only: floats between the partial outputs of consecutive slices
  int n_pass;           // 3 = 3xTF32, 1 = single-pass TF32
  int hi_chunks;        // accumulators the hi*hi term is spread over (0 = default, see gemm_tc.cu)
  int a_row0, b_row0;   // row origin of this GEMM inside the A / B tensor maps (e.g. t * Bp for tapes)
  int c_row0, x_row0;   // row origin of the output tile inside maps c0/c1, of the combined tile inside x0/x1
  int ldc;              // leading dimension (floats) of the output (used by colsum_part only)
  int act;              // HDPO_ACT_* of the epilogue
  int pdl_late;         // adjoint epilogues: let the dependent kernel start once this CTA's accumulators are complete
                        // instead of right after its dependency wait (set when chunk streams compete for the SMs)
  const float* bias;    // EPI_FWD_*
  TraceRef trace;       // optional per-CTA trace records (hdpo_debug_set_trace); tag set by the caller
  long long* dbg_clock; // optional: 8 clock64 stamps per CTA (tools/gemm_timeline.py)
  float* colsum_part;   // EPI_DGRAD_HIDDEN, optional: [rows / 32][ldc] column sums of each 32-row block of the output
                        // (bias-gradient partials; the row index includes a_row0, i.e. it follows the gz tape)
};

// Tensor maps of one GEMM launch. Operands: box = [128 | BN rows][32 floats] (K-major) or [32 k][32 mn] (MN-major).
// Outputs and the tiles an epilogue combines with: box = [kBoxRowsC rows][32 floats], 128-byte swizzle.
//   c0: output (hi half for EPI_FWD_HIDDEN / EPI_DGRAD_HIDDEN, full fp32 otherwise)   c1: lo half
//   x0 / x1: EPI_DGRAD_HIDDEN saved layer output h = x0 + x1 (act' is evaluated from it); EPI_DGRAD_ACCUM: x0 = the
//            array the product is added to (normally the same array as c0)
struct GemmTcMaps {
  CUtensorMap a_hi, a_lo, b_hi, b_lo, c0, c1, x0, x1;
};

// one 2-D fp32 tensor map over a row-major [rows][ld] array, box = [box_rows][32 floats]; 128B swizzle for K-major
// operands, 128B swizzle with 32-byte atoms for the MN-major (weight-gradient) operands
int make_tensor_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    bool mn_major = false);

// weight-gradient form: D[M,N] (per K slice) = sum_k A[k][m] * B[k][n] with A = [K][M], B = [K][N] row-major arrays
// (tensor maps with box_rows = 32); partial slice z is written at c_full + z * c_slice.
// (the partial slices form one [n_slices * M][N] array behind map c0)
int gemm_wgrad(const GemmTcMaps& tm, const GemmTcArgs& g, int bn, void* stream);

int gemm(const GemmTcMaps& tm, const GemmTcArgs& g, int epi, int bn, void* stream);

constexpr int kBoxRowsA = 128;
constexpr int kBoxRowsC = 32;
inline int pick_bn(int N) { return (N % 128 == 0) ? 128 : 64; }
// `bn` selector of gemm(): 64 / 128 = single-CTA tiles 128 x bn; kBnPair = CTA-pair (cta_group::2) tiles 256 x 128.
constexpr int kBnPair = 1128;
// kBnMulti + 64 | 128 = multi-tile CTA-pair form (wide_persist.cu, gemm_multi): 256 x 64 | 128 tiles, several per pair
constexpr int kBnMulti = 2000;
// two CTAs per SM (4 x 64 TMEM columns and a 2-stage ring each): kBnPair64 = CTA-pair tiles 256 x 64, kBn64x2 = single-CTA
// tiles 128 x 64
constexpr int kBnPair64 = 1064;
constexpr int kBn64x2 = 1065;
inline int b_box_rows(int bn) {  // rows of the B operand's TMA box
  if (bn == kBnPair) return 64;
  if (bn == kBnPair64) return 32;
  if (bn == kBn64x2) return 64;
  return bn > kBnMulti ? (bn - kBnMulti) / 2 : bn;
}
// the pair form needs 256-row and 128-column tiles; HDPO_TC_PAIR=0 disables it (A/B comparison on the GPU box)
int pair_enabled();
inline int pick_bn_pair(int M, int N) { return (pair_enabled() && M % 256 == 0 && N % 128 == 0) ? kBnPair : pick_bn(N); }

}  // namespace tc
}  // namespace hdpo
#endif
