// mma32.cuh - warp-level tensor-core forms of the 32-wide policy layers (store net of the SymmetryAware heads).
//
// One warp owns 32 (x MT/2) rows that live as shared-memory ROWS (one row per lane in the SIMT code around it). The
// three GEMM-shaped steps of a hidden layer run on the tensor cores with mma.sync.m16n8k8 (tf32 inputs, fp32
// accumulate) and the 3xTF32 split (lo*hi + hi*lo + hi*hi: fp32-grade, measured 5.6e-7 abs at |y| ~ 1 against 3.6e-7
// of the FFMA form):
//   layer  : out[r][n] = bias[n] + sum_k in[r][k] W[n][k]           M = rows, N = 32, K = 8 KT
//   dgrad  : same form with the transposed weight copy and a fused "* act'(h)" epilogue
//   wgrad  : acc[n][k] += sum_r G[r][n] X[r][k]                       M = 32, N = 8 NT, K = 32 rows
// Fragments come straight from the row arrays (strides 4*odd floats: conflict-free or 2-way), the weights from
// pre-split (hi, lo) copies W[n][k] with the same kind of stride. tools/mma32_bench.cu: 2.0x the packed-FFMA2 form
// per layer at 16 warps per SM (legacy HMMA path: no TMEM / mbarrier round trip, which a 32 x 32 x 32 product per
// warp could not amortise). Not available to the host-thread emulator: callers keep the SIMT form under HDPO_EMU.
#pragma once

#include "hdpo_platform.cuh"

#ifndef HDPO_EMU
namespace hdpo {
namespace mma32 {

__device__ __forceinline__ unsigned tf32_bits(float x) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void split(float v, unsigned& hi, unsigned& lo) {
  hi = tf32_bits(v);
  lo = tf32_bits(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// c[mt][nt] = A-rows x W^T for 16 MT rows and 32 outputs; W = (Whi, Wlo)[n][k] with row stride ws; K = 8 KT.
// Fragment layout of c[mt][nt][i]: row 16 mt + (lane >> 2) + 8 (i >> 1), column 8 nt + 2 (lane & 3) + (i & 1).
template <int MT>
__device__ __forceinline__ void product(const float* __restrict__ Whi, const float* __restrict__ Wlo, int ws,
                                        const float* __restrict__ in_rows, int in_stride, int KT, int lane,
                                        float (&c)[MT][4][4]) {
  const int g = lane >> 2, t = lane & 3;
  for (int kt = 0; kt < KT; ++kt) {
    unsigned ahi[MT][4], alo[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const float* r0 = in_rows + (16 * mt + g) * in_stride + 8 * kt + t;
      split(r0[0], ahi[mt][0], alo[mt][0]);
      split(r0[8 * in_stride], ahi[mt][1], alo[mt][1]);
      split(r0[4], ahi[mt][2], alo[mt][2]);
      split(r0[8 * in_stride + 4], ahi[mt][3], alo[mt][3]);
    }
    unsigned bh[4][2], bl[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* wh = Whi + (8 * nt + g) * ws + 8 * kt + t;
      const float* wl = Wlo + (8 * nt + g) * ws + 8 * kt + t;
      bh[nt][0] = __float_as_uint(wh[0]);
      bh[nt][1] = __float_as_uint(wh[4]);
      bl[nt][0] = __float_as_uint(wl[0]);
      bl[nt][1] = __float_as_uint(wl[4]);
    }
    // Issue order: one pass of the split over ALL accumulators before the next pass, so that consecutive HMMAs never
    // depend on each other (a warp issues in order: three back-to-back products into one accumulator expose the
    // tensor pipe's latency three times per tile; measured at 8 warps per SM).
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], alo[mt], bh[nt][0], bh[nt][1]);  // small terms first
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], ahi[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], ahi[mt], bh[nt][0], bh[nt][1]);
  }
}

// out[r][n] = bias[n] + sum_k in[r][k] W[n][k] for the 16 MT rows of this warp (pre-activations; rows may alias `in`
// only if the caller separates the phases with __syncwarp, which this function does before storing)
template <int MT>
__device__ __forceinline__ void layer(const float* __restrict__ Whi, const float* __restrict__ Wlo, int ws,
                                      const float* __restrict__ bias, const float* __restrict__ in_rows, int in_stride,
                                      int KT, float* __restrict__ out_rows, int out_stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float c[MT][4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 bv = *reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      c[mt][nt][0] = bv.x;
      c[mt][nt][1] = bv.y;
      c[mt][nt][2] = bv.x;
      c[mt][nt][3] = bv.y;
    }
  }
  product<MT>(Whi, Wlo, ws, in_rows, in_stride, KT, lane, c);
  __syncwarp();  // every lane has read its input fragments before rows are overwritten (in-place forward)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = out_rows + (16 * mt + g) * out_stride + 8 * nt + 2 * t;
      *reinterpret_cast<float2*>(o) = make_float2(c[mt][nt][0], c[mt][nt][1]);
      *reinterpret_cast<float2*>(o + 8 * out_stride) = make_float2(c[mt][nt][2], c[mt][nt][3]);
    }
  __syncwarp();
}

// In-place adjoint step through a hidden layer: h_rows[r][k] <- (sum_n gz[r][n] W[n][k]) * dact(h_rows[r][k]).
// (Wk_hi, Wk_lo)[k][n] = the TRANSPOSED weight copy; DACT(y) = activation derivative from the layer output.
template <class DACT>
__device__ __forceinline__ void dgrad_inplace(const float* __restrict__ Wk_hi, const float* __restrict__ Wk_lo, int ws,
                                              const float* __restrict__ gz_rows, int gz_stride,
                                              float* __restrict__ h_rows, int h_stride, int lane, DACT dact) {
  const int g = lane >> 2, t = lane & 3;
  float c[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
  product<2>(Wk_hi, Wk_lo, ws, gz_rows, gz_stride, 4, lane, c);
  __syncwarp();  // nobody still reads the h rows as an operand (they are overwritten below)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = h_rows + (16 * mt + g) * h_stride + 8 * nt + 2 * t;
      const float2 h0 = *reinterpret_cast<const float2*>(o);
      const float2 h1 = *reinterpret_cast<const float2*>(o + 8 * h_stride);
      *reinterpret_cast<float2*>(o) = make_float2(c[mt][nt][0] * dact(h0.x), c[mt][nt][1] * dact(h0.y));
      *reinterpret_cast<float2*>(o + 8 * h_stride) = make_float2(c[mt][nt][2] * dact(h1.x), c[mt][nt][3] * dact(h1.y));
    }
  __syncwarp();
}

// acc[mt][nt][i] += sum over the warp's 32 rows r of G[r][n] X[r][k], n = 16 mt + (lane >> 2) + 8 (i >> 1),
// k = 8 nt + 2 (lane & 3) + (i & 1). Both operands are activations: split on the fly. The tensor core's accumulation
// of the 12 products is kept out of the long-lived sum: it goes to a zeroed fragment that is then added with
// round-to-nearest (the accumulators live for a whole launch).
template <int NT>
__device__ __forceinline__ void wgrad(const float* __restrict__ G, int gs, const float* __restrict__ X, int xs, int lane,
                                      float (&acc)[2][NT][4]) {
  const int g = lane >> 2, t = lane & 3;
  float c[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
#pragma unroll 1
  for (int kt = 0; kt < 4; ++kt) {  // 8 rows per step
    unsigned ahi[2][4], alo[2][4];
    const float* g0 = G + (8 * kt + t) * gs + g;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {  // A[m = n][k = r] = G[r][n]
      split(g0[16 * mt], ahi[mt][0], alo[mt][0]);
      split(g0[16 * mt + 8], ahi[mt][1], alo[mt][1]);
      split(g0[4 * gs + 16 * mt], ahi[mt][2], alo[mt][2]);
      split(g0[4 * gs + 16 * mt + 8], ahi[mt][3], alo[mt][3]);
    }
    const float* x0 = X + (8 * kt + t) * xs + g;
    unsigned bh[NT][2], bl[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {  // B[k = r][n = k] = X[r][k]
      split(x0[8 * nt], bh[nt][0], bl[nt][0]);
      split(x0[4 * xs + 8 * nt], bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], alo[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], ahi[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt][nt], ahi[mt], bh[nt][0], bh[nt][1]);
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] += c[mt][nt][i];
}

// ---- variants with ONE fp32 weight copy W[n][k] (row stride ws), split into (hi, lo) on the fly: 3 ALU operations
// per B element instead of a second and third copy in shared memory (the small-net adjoint keeps two CTAs per SM).
// TRANSPOSED = false: c[r][n] += sum_k A[r][k] W[n][k]   (layer);  true: c[r][k] += sum_n A[r][n] W[n][k]   (dgrad)
template <int MT, bool TRANSPOSED, int NT = 4>
__device__ __forceinline__ void product_f32(const float* __restrict__ W, int ws, const float* __restrict__ in_rows,
                                            int in_stride, int KT, int lane, float (&c)[MT][NT][4]) {
  const int g = lane >> 2, t = lane & 3;
  for (int kt = 0; kt < KT; ++kt) {
    unsigned ahi[MT][4], alo[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const float* r0 = in_rows + (16 * mt + g) * in_stride + 8 * kt + t;
      split(r0[0], ahi[mt][0], alo[mt][0]);
      split(r0[8 * in_stride], ahi[mt][1], alo[mt][1]);
      split(r0[4], ahi[mt][2], alo[mt][2]);
      split(r0[8 * in_stride + 4], ahi[mt][3], alo[mt][3]);
    }
    unsigned bh[NT][2], bl[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      // B fragment element (contraction index 8 kt + t (+4), output index 8 nt + g)
      const float* w = TRANSPOSED ? W + (8 * kt + t) * ws + 8 * nt + g : W + (8 * nt + g) * ws + 8 * kt + t;
      split(w[0], bh[nt][0], bl[nt][0]);
      split(w[TRANSPOSED ? 4 * ws : 4], bh[nt][1], bl[nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], alo[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], ahi[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(c[mt][nt], ahi[mt], bh[nt][0], bh[nt][1]);
  }
}

template <int MT>
__device__ __forceinline__ void layer_f32(const float* __restrict__ W, int ws, const float* __restrict__ bias,
                                          const float* __restrict__ in_rows, int in_stride, int KT,
                                          float* __restrict__ out_rows, int out_stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float c[MT][4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 bv = *reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      c[mt][nt][0] = bv.x;
      c[mt][nt][1] = bv.y;
      c[mt][nt][2] = bv.x;
      c[mt][nt][3] = bv.y;
    }
  }
  product_f32<MT, false>(W, ws, in_rows, in_stride, KT, lane, c);
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = out_rows + (16 * mt + g) * out_stride + 8 * nt + 2 * t;
      *reinterpret_cast<float2*>(o) = make_float2(c[mt][nt][0], c[mt][nt][1]);
      *reinterpret_cast<float2*>(o + 8 * out_stride) = make_float2(c[mt][nt][2], c[mt][nt][3]);
    }
  __syncwarp();
}

template <class DACT>
__device__ __forceinline__ void dgrad_inplace_f32(const float* __restrict__ W, int ws, const float* __restrict__ gz_rows,
                                                  int gz_stride, float* __restrict__ h_rows, int h_stride, int lane,
                                                  DACT dact) {
  const int g = lane >> 2, t = lane & 3;
  float c[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
  product_f32<2, true>(W, ws, gz_rows, gz_stride, 4, lane, c);
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* o = h_rows + (16 * mt + g) * h_stride + 8 * nt + 2 * t;
      const float2 h0 = *reinterpret_cast<const float2*>(o);
      const float2 h1 = *reinterpret_cast<const float2*>(o + 8 * h_stride);
      *reinterpret_cast<float2*>(o) = make_float2(c[mt][nt][0] * dact(h0.x), c[mt][nt][1] * dact(h0.y));
      *reinterpret_cast<float2*>(o + 8 * h_stride) = make_float2(c[mt][nt][2] * dact(h1.x), c[mt][nt][3] * dact(h1.y));
    }
  __syncwarp();
}

// rows[r][k] += sum_n gz[r][n] W[n][k] for k < 8 NT (first-layer dgrad into the state-adjoint rows; no activation)
template <int NT>
__device__ __forceinline__ void dgrad_accum_f32(const float* __restrict__ W, int ws, const float* __restrict__ gz_rows,
                                                int gz_stride, float* __restrict__ rows, int row_stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float c[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
  product_f32<2, true, NT>(W, ws, gz_rows, gz_stride, 4, lane, c);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* o = rows + (16 * mt + g) * row_stride + 8 * nt + 2 * t;
      float2 v0 = *reinterpret_cast<const float2*>(o);
      float2 v1 = *reinterpret_cast<const float2*>(o + 8 * row_stride);
      v0.x += c[mt][nt][0];
      v0.y += c[mt][nt][1];
      v1.x += c[mt][nt][2];
      v1.y += c[mt][nt][3];
      *reinterpret_cast<float2*>(o) = v0;
      *reinterpret_cast<float2*>(o + 8 * row_stride) = v1;
    }
  __syncwarp();
}

}  // namespace mma32
}  // namespace hdpo
#endif  // HDPO_EMU
