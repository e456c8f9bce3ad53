// hdpo_math.cuh - activations and their derivatives with torch semantics (neural_networks.py:36-43).
// Precise libm-grade device functions on purpose (expm1f/log1pf/expf, no --use_fast_math): the parity bar is
// 1e-5 relative against a true-fp32 reference.
#pragma once

#include <type_traits>

#include "hdpo_platform.cuh"
#include "../../include/hdpo_b200.h"

namespace hdpo {

// expm1(x) for x <= 0 in ~10 instructions at <= 2 ulp: degree-7 Taylor/Horner near zero (|x| <= 0.35, truncation
// 1.6e-8 relative), expf(x) - 1 below that (result magnitude >= 0.29, so the absolute error of expf dominates).
// libm's expm1f is ~25 instructions; with 96 ELUs per scenario-period that cost as much as the FFMAs themselves.
__device__ __forceinline__ float expm1_nonpos(float x) {
  float p = fmaf(x, 1.f / 5040.f, 1.f / 720.f);
  p = fmaf(x, p, 1.f / 120.f);
  p = fmaf(x, p, 1.f / 24.f);
  p = fmaf(x, p, 1.f / 6.f);
  p = fmaf(x, p, 0.5f);
  p = fmaf(x, p, 1.f);
  p = x * p;
#ifdef HDPO_EMU
  return x > -0.35f ? p : expf(x) - 1.f;
#else
  // below -0.35 the result is >= 0.29 in magnitude, so MUFU.EX2's 2^-22 relative error on exp(x) stays below 1e-6
  // relative on the result: __expf (FMUL + MUFU) instead of the ~8-instruction precise expf
  return x > -0.35f ? p : __expf(x) - 1.f;
#endif
}

// Branch-free ELU over a small register array (same arithmetic as elu_f). In a tile epilogue every lane holds 16
// unrelated values: the per-element `x > 0 ? ... : ...` of elu_f compiles to one divergent region per element with a
// serial 7-FMA Horner chain inside (measured: 5.8 us per 128 x 128 tile, as long as the tile's whole mainloop);
// evaluating both branches for all elements lets the chains of different elements overlap.
#ifndef HDPO_EMU
// packed fp32 pairs (Blackwell f32x2 forms -> SASS FFMA2 / FADD2 / FMUL2): two independent round-to-nearest operations
// per issued instruction, bit-identical to the scalar ones
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2_rn(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2_rn(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2_rn(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#endif

template <int N, bool PACKED = true>
__device__ __forceinline__ void elu_inplace(float (&v)[N]) {
#ifndef HDPO_EMU
  if constexpr (PACKED && N % 2 == 0) {
    // same arithmetic as the scalar form below, the Horner steps / products / the "- 1" on packed pairs
    unsigned long long xm[N / 2], p[N / 2], e[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; ++j) xm[j] = pack2(fminf(v[2 * j], 0.f), fminf(v[2 * j + 1], 0.f));
    const unsigned long long c7 = pack2(1.f / 5040.f, 1.f / 5040.f), c6 = pack2(1.f / 720.f, 1.f / 720.f),
                             c5 = pack2(1.f / 120.f, 1.f / 120.f), c4 = pack2(1.f / 24.f, 1.f / 24.f),
                             c3 = pack2(1.f / 6.f, 1.f / 6.f), c2 = pack2(0.5f, 0.5f), c1 = pack2(1.f, 1.f),
                             m1 = pack2(-1.f, -1.f);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], c7, c6);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], p[j], c5);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], p[j], c4);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], p[j], c3);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], p[j], c2);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = fma2_rn(xm[j], p[j], c1);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) p[j] = mul2_rn(xm[j], p[j]);
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      float x0, x1;
      unpack2(xm[j], x0, x1);
      e[j] = add2_rn(pack2(__expf(x0), __expf(x1)), m1);
    }
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      float x0, x1, p0, p1, e0, e1;
      unpack2(xm[j], x0, x1);
      unpack2(p[j], p0, p1);
      unpack2(e[j], e0, e1);
      const float r0 = x0 > -0.35f ? p0 : e0, r1 = x1 > -0.35f ? p1 : e1;
      v[2 * j] = v[2 * j] > 0.f ? v[2 * j] : r0;
      v[2 * j + 1] = v[2 * j + 1] > 0.f ? v[2 * j + 1] : r1;
    }
  } else
#endif
  {
  float xm[N], p[N], e[N];
#pragma unroll
  for (int j = 0; j < N; ++j) xm[j] = fminf(v[j], 0.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], 1.f / 5040.f, 1.f / 720.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], p[j], 1.f / 120.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], p[j], 1.f / 24.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], p[j], 1.f / 6.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], p[j], 0.5f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = fmaf(xm[j], p[j], 1.f);
#pragma unroll
  for (int j = 0; j < N; ++j) p[j] = xm[j] * p[j];
#pragma unroll
  for (int j = 0; j < N; ++j) {
#ifdef HDPO_EMU
    e[j] = expf(xm[j]) - 1.f;
#else
    e[j] = __expf(xm[j]) - 1.f;
#endif
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float r = xm[j] > -0.35f ? p[j] : e[j];
    v[j] = v[j] > 0.f ? v[j] : r;
  }
  }
}

// nn.ELU(alpha=1): x > 0 ? x : expm1(x)
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1_nonpos(x); }
// derivative expressed with the OUTPUT y: 1 for x > 0, exp(x) = y + 1 for x <= 0 (y == 0 at x == 0 -> 1)
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }

// nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad(float x) {
  if (x > 20.f) return 1.f;
  float z = expf(x);
  return z / (z + 1.f);
}

__device__ __forceinline__ float sigmoid_f(float x) {
  float e = expf(-fabsf(x));
  return x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
}

// generic activation by id; `y` = output, used for derivatives where torch does the same
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case HDPO_ACT_ELU: return elu_f(x);
    case HDPO_ACT_RELU: return fmaxf(x, 0.f);
    case HDPO_ACT_TANH: return tanhf(x);
    case HDPO_ACT_SIGMOID: return sigmoid_f(x);
    case HDPO_ACT_SOFTPLUS: return softplus_f(x);
    default: return x;
  }
}
// derivative given pre-activation x and output y
__device__ __forceinline__ float act_grad(int act, float x, float y) {
  switch (act) {
    case HDPO_ACT_ELU: return x > 0.f ? 1.f : y + 1.f;
    case HDPO_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case HDPO_ACT_TANH: return 1.f - y * y;
    case HDPO_ACT_SIGMOID: return y * (1.f - y);
    case HDPO_ACT_SOFTPLUS: return softplus_grad(x);
    default: return 1.f;
  }
}
// derivative from the output alone (exact for elu/relu/tanh/sigmoid/none; softplus: 1 - exp(-y))
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case HDPO_ACT_ELU: return elu_grad_from_out(y);
    case HDPO_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case HDPO_ACT_TANH: return 1.f - y * y;
    case HDPO_ACT_SIGMOID: return y * (1.f - y);
    case HDPO_ACT_SOFTPLUS: return y > 20.f ? 1.f : -expm1f(-y);
    default: return 1.f;
  }
}

// Compile-time activation variants + a row-level dispatcher. Applying the runtime switch PER ELEMENT inside unrolled
// loops inlines all six activations (tanhf / log1pf / division slow paths) 32x per row and blew the kernels up to
// 300 KB of SASS (instruction-cache thrash, ncu "no_instruction" stalls); dispatching once per row keeps the hot
// loop to the one activation actually in use.
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float x) {
  if (ACT == HDPO_ACT_ELU) return elu_f(x);
  if (ACT == HDPO_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == HDPO_ACT_TANH) return tanhf(x);
  if (ACT == HDPO_ACT_SIGMOID) return sigmoid_f(x);
  if (ACT == HDPO_ACT_SOFTPLUS) return softplus_f(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_grad_out_t(float y) {
  if (ACT == HDPO_ACT_ELU) return elu_grad_from_out(y);
  if (ACT == HDPO_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (ACT == HDPO_ACT_TANH) return 1.f - y * y;
  if (ACT == HDPO_ACT_SIGMOID) return y * (1.f - y);
  if (ACT == HDPO_ACT_SOFTPLUS) return y > 20.f ? 1.f : -expm1f(-y);
  return 1.f;
}
template <class F>
__device__ __forceinline__ void dispatch_act(int act, F&& f) {
  switch (act) {
    case HDPO_ACT_ELU: f(std::integral_constant<int, HDPO_ACT_ELU>{}); break;
    case HDPO_ACT_RELU: f(std::integral_constant<int, HDPO_ACT_RELU>{}); break;
    case HDPO_ACT_TANH: f(std::integral_constant<int, HDPO_ACT_TANH>{}); break;
    case HDPO_ACT_SIGMOID: f(std::integral_constant<int, HDPO_ACT_SIGMOID>{}); break;
    case HDPO_ACT_SOFTPLUS: f(std::integral_constant<int, HDPO_ACT_SOFTPLUS>{}); break;
    default: f(std::integral_constant<int, HDPO_ACT_NONE>{}); break;
  }
}

// Packed fp32 FMA (Blackwell: fma.rn.f32x2 -> SASS FFMA2): two independent round-to-nearest FMAs per issued
// instruction, bit-identical to two fmaf() calls. The small-net kernels are issue-bound (FMA ~50 % of the issued
// instructions), so halving the FMA instruction count is a direct win.
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
#ifdef HDPO_EMU
  d.x = fmaf(a.x, b.x, d.x);
  d.y = fmaf(a.y, b.y, d.y);
#else
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
#endif
}

// clip(x, min=0) and its torch sub-gradient (1 for x >= 0, incl. x == 0)
__device__ __forceinline__ float relu0(float x) { return fmaxf(x, 0.f); }
__device__ __forceinline__ float ge0(float x) { return x >= 0.f ? 1.f : 0.f; }
__device__ __forceinline__ float le0(float x) { return x <= 0.f ? 1.f : 0.f; }

}  // namespace hdpo
