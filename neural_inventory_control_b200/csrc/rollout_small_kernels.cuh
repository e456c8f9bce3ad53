// rollout_small_kernels.cuh - device code of the fused small-net rollout (see rollout_small.cu for the design notes).
// Included by rollout_small.cu (forward + host wrappers) and by the rollout_small_bwd_*.cu instantiation units
// (the adjoint kernel is templated on <policy arch, input tile width, hidden depth>; the instantiations are spread
// over several translation units so that nvcc builds them in parallel).
#pragma once

#include "mma32.cuh"
#include "rollout_small.cuh"

namespace hdpo {
namespace small {

constexpr int kWarpsPerCta = 4;
constexpr int kMaxPartialRows = 4096;

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------

struct Statics {
  float h, p, lt, ltw, hw, edge;
  float lte[kMaxE], he[kMaxE];
};

template <int ARCH>
__device__ __forceinline__ void load_statics(const Cfg& c, const HdpoStatics& st, int b, Statics& s) {
  s.h = st.holding_costs[b];
  s.p = st.underage_costs[b];
  s.lt = st.lead_times[b];
  s.ltw = s.hw = s.edge = 0.f;
  if (ARCH != HDPO_ARCH_VANILLA_ONE_STORE && c.W > 0) {
    s.ltw = st.warehouse_lead_times[b];
    s.hw = st.warehouse_holding_costs[b];
    if (c.has_edge) s.edge = st.warehouse_edge_costs[b];
  }
#pragma unroll
  for (int e = 0; e < kMaxE; ++e) {
    const bool on = ARCH != HDPO_ARCH_VANILLA_ONE_STORE && e < c.E;
    s.lte[e] = on ? st.echelon_lead_times[b * c.E + e] : 1.f;
    s.he[e] = on ? st.echelon_holding_costs[b * c.E + e] : 0.f;
  }
}

// stage the parameter vector into the shared weight block: Wt[k][n] (transposed, zero padded), biases, Wo[o][k]
static __device__ void stage_weights(const Cfg& c, const float* __restrict__ params, float* __restrict__ Ws,
                                     bool bwd = false) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n_floats = bwd ? c.s_total_bwd : c.s_total;
  for (int i = tid; i < n_floats; i += nt) Ws[i] = 0.f;
  __syncthreads();
  {  // layer 0: W0 [w1][IN]
    const int n_out = c.w[1], n_in = c.IN;
    for (int i = tid; i < n_out * n_in; i += nt) {
      int n = i / n_in, k = i % n_in;
      Ws[c.s_wt0 + k * H + n] = params[c.gw[0] + i];
      if (bwd && c.tc) Ws[c.s_w0n + n * c.XSb + k] = params[c.gw[0] + i];  // W0[n][k] for the mma fragments
    }
    for (int i = tid; i < n_out; i += nt) Ws[c.s_b0 + i] = params[c.gb[0] + i];
  }
  for (int l = 0; l < c.NHH; ++l) {
    const int n_out = c.w[l + 2], n_in = c.w[l + 1];
    for (int i = tid; i < n_out * n_in; i += nt) {
      int n = i / n_in, k = i % n_in;
      Ws[c.s_wth[l] + k * H + n] = params[c.gw[l + 1] + i];
      if (bwd && c.tc) Ws[c.s_wn[l] + n * HS + k] = params[c.gw[l + 1] + i];  // W[n][k] for the mma fragments
    }
    for (int i = tid; i < n_out; i += nt) Ws[c.s_bh[l] + i] = params[c.gb[l + 1] + i];
  }
  {  // output layer: Wo [OUT][w_last] kept row-major [o][k]
    const int n_out = c.OUT, n_in = c.w[c.NHH + 1];
    for (int i = tid; i < n_out * n_in; i += nt) {
      int o = i / n_in, k = i % n_in;
      Ws[c.s_wo + o * H + k] = params[c.gw[c.NHH + 1] + i];
    }
    for (int i = tid; i < n_out; i += nt) Ws[c.s_bo + i] = params[c.gb[c.NHH + 1] + i];
  }
  __syncthreads();
}

__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// acc[j][n] = bias[n] + sum_k Wt[k][n] * in_j[k]   for NS scenario rows handled by this lane.
// Accumulators are float2 pairs (n even, n odd) fed by packed FFMA2: per k and scenario 16 FFMA2 + one (x,x) pack.
template <int NS>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ Wt, const float* __restrict__ bias, int K4,
                                          const float* const (&in)[NS], float2 (&acc)[NS][H / 2]) {
#pragma unroll
  for (int n4 = 0; n4 < H / 4; ++n4) {
    const float4 bv = reinterpret_cast<const float4*>(bias)[n4];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      acc[j][2 * n4 + 0] = make_float2(bv.x, bv.y);
      acc[j][2 * n4 + 1] = make_float2(bv.z, bv.w);
    }
  }
  for (int k4 = 0; k4 < K4; ++k4) {
    float4 xv[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) xv[j] = reinterpret_cast<const float4*>(in[j])[k4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4* w = reinterpret_cast<const float4*>(Wt + (4 * k4 + kk) * H);
      float2 xx[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const float xk = f4c(xv[j], kk);
        xx[j] = make_float2(xk, xk);
      }
#pragma unroll
      for (int n4 = 0; n4 < H / 4; ++n4) {
        const float4 wv = w[n4];
        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          ffma2(acc[j][2 * n4 + 0], w01, xx[j]);
          ffma2(acc[j][2 * n4 + 1], w23, xx[j]);
        }
      }
    }
  }
}

// apply the hidden activation and store the row
__device__ __forceinline__ void act_store_row(int act, const float2 (&acc)[H / 2], float* __restrict__ row) {
  dispatch_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll
    for (int n4 = 0; n4 < H / 4; ++n4) {
      float4 v;
      v.x = act_fwd_t<ACT>(acc[2 * n4 + 0].x);
      v.y = act_fwd_t<ACT>(acc[2 * n4 + 0].y);
      v.z = act_fwd_t<ACT>(acc[2 * n4 + 1].x);
      v.w = act_fwd_t<ACT>(acc[2 * n4 + 1].y);
      reinterpret_cast<float4*>(row)[n4] = v;
    }
  });
}

// y[o] = bo[o] + sum_k Wo[o][k] * h[k]
__device__ __forceinline__ void out_layer_fwd(const Cfg& c, const float* __restrict__ Ws, const float* __restrict__ hrow,
                                              float (&y)[kMaxOut]) {
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    if (o < c.OUT) {
      const float4* w = reinterpret_cast<const float4*>(Ws + c.s_wo + o * H);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 wv = w[k4];
        const float4 hv = reinterpret_cast<const float4*>(hrow)[k4];
        s0 = fmaf(wv.x, hv.x, s0);
        s1 = fmaf(wv.y, hv.y, s1);
        s2 = fmaf(wv.z, hv.z, s2);
        s3 = fmaf(wv.w, hv.w, s3);
      }
      y[o] = Ws[c.s_bo + o] + ((s0 + s1) + (s2 + s3));
    } else {
      y[o] = 0.f;
    }
  }
}

// full MLP forward for NS rows; hidden activations of every layer are left in hb[l] rows (l = 0..NHH)
template <int NS>
__device__ __forceinline__ void mlp_fwd(const Cfg& c, const float* __restrict__ Ws, const float* const (&xrow)[NS],
                                        float* const (&hrow)[NS], int h_layer_stride, float (&y)[NS][kMaxOut]) {
  float2 acc[NS][H / 2];
  layer_fwd<NS>(Ws + c.s_wt0, Ws + c.s_b0, c.IN4 / 4, xrow, acc);
#pragma unroll
  for (int j = 0; j < NS; ++j) act_store_row(c.hidden_act, acc[j], hrow[j]);
  for (int l = 0; l < c.NHH; ++l) {
    const float* in[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) in[j] = hrow[j] + l * h_layer_stride;
    layer_fwd<NS>(Ws + c.s_wth[l], Ws + c.s_bh[l], H / 4, in, acc);
#pragma unroll
    for (int j = 0; j < NS; ++j) act_store_row(c.hidden_act, acc[j], hrow[j] + (l + 1) * h_layer_stride);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) out_layer_fwd(c, Ws, hrow[j] + c.NHH * h_layer_stride, y[j]);
}

// hidden activation applied in place to this lane's row (rolled loop: instruction-cache footprint)
__device__ __forceinline__ void act_row_inplace(int act, float* __restrict__ row) {
  dispatch_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll 1
    for (int n4 = 0; n4 < H / 4; ++n4) {
      const float4 t = reinterpret_cast<float4*>(row)[n4];
      float v[4] = {t.x, t.y, t.z, t.w};
      if (ACT == HDPO_ACT_ELU) {
        elu_inplace(v);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = act_fwd_t<ACT>(v[i]);
      }
      reinterpret_cast<float4*>(row)[n4] = make_float4(v[0], v[1], v[2], v[3]);
    }
  });
}

#ifndef HDPO_EMU
// Recompute of the adjoint kernel in tensor-core mode: first layer (inputs zero-padded to K0 = a multiple of 8 in the
// state rows) and HxH layers as warp-level mma.sync products over the warp's 32 rows (mma32.cuh), output layer in the
// FFMA form. Xw / Hb = bases of the warp's state rows [32][XS] and activation rows [NHH + 1][32][HS].
__device__ __forceinline__ void mlp_fwd_tc(const Cfg& c, const float* __restrict__ Ws, const float* __restrict__ Xw,
                                           float* __restrict__ Hb, int layer_stride, int lane, float (&y)[kMaxOut]) {
  float* hrow = Hb + lane * HS;
  __syncwarp();  // the state rows were written lane by lane
  mma32::layer_f32<2>(Ws + c.s_w0n, c.XSb, Ws + c.s_b0, Xw, c.XSb, c.K0 / 8, Hb, HS, lane);
  act_row_inplace(c.hidden_act, hrow);
  for (int l = 0; l < c.NHH; ++l) {
    __syncwarp();
    mma32::layer_f32<2>(Ws + c.s_wn[l], HS, Ws + c.s_bh[l], Hb + l * layer_stride, HS, H / 8, Hb + (l + 1) * layer_stride,
                        HS, lane);
    act_row_inplace(c.hidden_act, hrow + (l + 1) * layer_stride);
  }
  out_layer_fwd(c, Ws, hrow + c.NHH * layer_stride, y);
}
#endif

// ---- policy head (neural_networks.py:211-214 / 335-349): y -> allocations a[], plus what the adjoint needs
struct Head {
  float a[kMaxOut];      // one_store: a[0]; serial: a[e] echelons, a[E] warehouse, a[E+1] store
  float sg[kMaxOut];     // serial: sigmoid(y)
  float bound[kMaxOut];  // serial: upper bounds
  float pre;             // one_store: y + 1
};

template <int ARCH>
__device__ __forceinline__ void head_fwd(const Cfg& c, const float* __restrict__ x, const float (&y)[kMaxOut],
                                         Head& hd) {
  if (ARCH == HDPO_ARCH_VANILLA_ONE_STORE) {
    hd.pre = y[0] + 1.f;
    hd.a[0] = softplus_f(hd.pre);
  } else {
    const float* xw = x + c.L;
    const float* xe = xw + c.Lw;
#pragma unroll
    for (int i = 0; i < kMaxOut; ++i) {
      if (i < c.E + 2) {
        float bnd = (i == 0) ? c.wub : (i <= c.E ? xe[(i - 1) * c.Le] : xw[0]);
        hd.bound[i] = bnd;
        hd.sg[i] = sigmoid_f(y[i]);
        hd.a[i] = hd.sg[i] * bnd;
      }
    }
  }
  if (c.discrete) {
#pragma unroll
    for (int i = 0; i < kMaxOut; ++i)
      if (i < c.OUT) hd.a[i] = rintf(hd.a[i]);
  }
}

__device__ __forceinline__ void shift_left(float* __restrict__ x, int len, float post) {
  x[0] = post + x[1];
  for (int k = 1; k < len - 1; ++k) x[k] = x[k + 1];
  x[len - 1] = 0.f;
}
__device__ __forceinline__ void land(float* __restrict__ x, int len, float amount, float lead) {
  if (amount != 0.f) {
    int slot = static_cast<int>(lead) - 1;
    if (slot >= 0 && slot < len) x[slot] += amount;
  }
}

// one simulator period on the lane's state row, in place (environment.py:110-299). Returns the period cost.
template <int ARCH>
__device__ __forceinline__ float env_fwd(const Cfg& c, float* __restrict__ x, float d, const Head& hd,
                                         const Statics& s) {
  const float a_store = (ARCH == HDPO_ARCH_VANILLA_ONE_STORE) ? hd.a[0] : hd.a[c.E + 1];
  const float on_hand = x[0];
  const float raw = on_hand - d;
  float r;
  if (c.profit)
    r = -s.p * fminf(on_hand, d) + s.h * relu0(raw);
  else
    r = s.p * relu0(-raw) + s.h * relu0(raw);
  shift_left(x, c.L, c.lost ? relu0(raw) : raw);
  land(x, c.L, a_store, s.lt);
  if (ARCH != HDPO_ARCH_VANILLA_ONE_STORE && c.W > 0) {
    float* xw = x + c.L;
    const float a_wh = hd.a[c.E];
    const float raw_w = xw[0] - a_store;
    float cw = s.hw * relu0(raw_w);
    if (c.has_edge) cw += s.edge * a_wh;
    r += cw;
    shift_left(xw, c.Lw, raw_w);
    land(xw, c.Lw, a_wh, s.ltw);
    float re = 0.f;
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
      if (e < c.E) {
        float* xe = xw + c.Lw + e * c.Le;
        const float drawn = (e + 1 < c.E) ? hd.a[e + 1] : a_wh;
        const float raw_e = xe[0] - drawn;
        re += s.he[e] * relu0(raw_e);
        shift_left(xe, c.Le, raw_e);
        land(xe, c.Le, hd.a[e], s.lte[e]);
      }
    }
    if (c.E > 0) r += re;
  }
  return r;
}

__device__ __forceinline__ float gather_slot(const float* __restrict__ g, int len, float amount, float lead) {
  if (amount == 0.f) return 0.f;  // exact zeros were filtered before the put: no pipeline gradient
  int slot = static_cast<int>(lead) - 1;
  return (slot >= 0 && slot < len) ? g[slot] : 0.f;
}
__device__ __forceinline__ void shift_right(float* __restrict__ g, int len, float g0) {
  const float gn0 = g[0];
  for (int k = len - 1; k >= 2; --k) g[k] = g[k - 1];
  g[1] = gn0;
  g[0] = g0;
}

// adjoint of head_fwd + env_fwd for one lane. x = state BEFORE the period (read only), g = adjoint row: on entry
// the adjoint wrt the NEXT state, on exit the direct (non-MLP) part of the adjoint wrt x. gy = adjoint wrt MLP outputs.
template <int ARCH>
__device__ __forceinline__ void head_env_bwd(const Cfg& c, const float* __restrict__ x, float* __restrict__ g, float d,
                                             const Head& hd, const Statics& s, float rb, float (&gy)[kMaxOut]) {
  constexpr bool one = (ARCH == HDPO_ARCH_VANILLA_ONE_STORE);
  const float a_store = one ? hd.a[0] : hd.a[c.E + 1];
  const float on_hand = x[0];
  const float raw = on_hand - d;
  float ga_store = gather_slot(g, c.L, a_store, s.lt);
  float g0;
  if (c.profit) {
    const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
    g0 = rb * (-s.p * tie + s.h * ge0(raw));
  } else {
    g0 = rb * (-s.p * le0(raw) + s.h * ge0(raw));
  }
  g0 += c.lost ? g[0] * ge0(raw) : g[0];
  shift_right(g, c.L, g0);
  float ga[kMaxOut];
#pragma unroll
  for (int i = 0; i < kMaxOut; ++i) ga[i] = 0.f;
  if (!one && c.W > 0) {
    const float* xw = x + c.L;
    float* gw = g + c.L;
    const float a_wh = hd.a[c.E];
    const float raw_w = xw[0] - a_store;
    float ga_wh = gather_slot(gw, c.Lw, a_wh, s.ltw);
    const float g_raw_w = rb * s.hw * ge0(raw_w) + gw[0];
    shift_right(gw, c.Lw, g_raw_w);
    ga_store -= g_raw_w;
    if (c.has_edge) ga_wh += rb * s.edge;
    float g_raw_prev = 0.f;
#pragma unroll
    for (int e = 0; e < kMaxE; ++e) {
      if (e < c.E) {
        const float* xe = xw + c.Lw + e * c.Le;
        float* ge = gw + c.Lw + e * c.Le;
        const float drawn = (e + 1 < c.E) ? hd.a[e + 1] : a_wh;
        const float raw_e = xe[0] - drawn;
        float gae = gather_slot(ge, c.Le, hd.a[e], s.lte[e]);
        const float g_raw_e = rb * s.he[e] * ge0(raw_e) + ge[0];
        shift_right(ge, c.Le, g_raw_e);
        if (e >= 1) gae -= g_raw_prev;
        ga[e] = gae;
        g_raw_prev = g_raw_e;
      }
    }
    if (c.E > 0) ga_wh -= g_raw_prev;
    ga[c.E] = ga_wh;
    ga[c.E + 1] = ga_store;
  } else {
    ga[0] = ga_store;
  }
  if (one) {
    gy[0] = ga[0] * softplus_grad(hd.pre);
#pragma unroll
    for (int i = 1; i < kMaxOut; ++i) gy[i] = 0.f;
  } else {
    float* gw = g + c.L;
    float* gE = gw + c.Lw;
#pragma unroll
    for (int i = 0; i < kMaxOut; ++i) {
      if (i < c.E + 2) {
        const float sg = hd.sg[i];
        gy[i] = ga[i] * hd.bound[i] * sg * (1.f - sg);
        const float gb = ga[i] * sg;  // adjoint of the (non-detached) bound
        if (i >= 1 && i <= c.E) gE[(i - 1) * c.Le] += gb;
        if (i == c.E + 1) gw[0] += gb;
      } else {
        gy[i] = 0.f;
      }
    }
  }
}

// demand of (scenario b, period t) for S == 1
__device__ __forceinline__ float demand_at(const Cfg& c, const float* __restrict__ dem, int b, int t) {
  const int tt = t + c.period_shift;
  if (c.demand_layout == HDPO_DEMAND_TSB) return __ldg(dem + static_cast<int64_t>(tt) * c.B + b);
  return __ldg(dem + static_cast<int64_t>(b) * c.t_stride + tt);
}

// ------------------------------------------------------------------------------------------------------------
// K1: forward rollout
// ------------------------------------------------------------------------------------------------------------
template <int ARCH, int NS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
small_fwd_kernel(Cfg c, const float* __restrict__ params, const float* __restrict__ demands, HdpoStatics st,
                 HdpoState init, float* __restrict__ cost_b, float* __restrict__ report_b,
                 float* __restrict__ reward_tb, float* __restrict__ tape, HdpoState fin) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights(c, params, Ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = 32 * NS;
  float* Xw = smem + ((c.s_total + 3) & ~3) + warp * rows * (c.XS + HS);
  float* Hw = Xw + rows * c.XS;
  const int tile_scen = 32 * NS;
  const int n_tiles = ceil_div(c.B, tile_scen);
  const int wpc = blockDim.x >> 5;  // warps per CTA: 4 for big batches, 2 / 1 when there are few tiles (latency)
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;

  for (int tile = gwarp; tile < n_tiles; tile += nwarps) {
    int b[NS];
    bool valid[NS];
    float* xrow[NS];
    float* hrow[NS];
    Statics s[NS];
    float cost[NS], rep[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      int bb = tile * tile_scen + j * 32 + lane;
      valid[j] = bb < c.B;
      b[j] = valid[j] ? bb : c.B - 1;
      xrow[j] = Xw + (j * 32 + lane) * c.XS;
      hrow[j] = Hw + (j * 32 + lane) * HS;
      load_statics<ARCH>(c, st, b[j], s[j]);
      cost[j] = rep[j] = 0.f;
      // initial state row: [store L | warehouse Lw | echelons E*Le], zero padded to IN4
      for (int k = 0; k < c.L; ++k) xrow[j][k] = init.store[static_cast<int64_t>(b[j]) * c.L + k];
      for (int k = 0; k < c.Lw; ++k) xrow[j][c.L + k] = init.warehouse[static_cast<int64_t>(b[j]) * c.Lw + k];
      for (int k = 0; k < c.E * c.Le; ++k)
        xrow[j][c.L + c.Lw + k] = init.echelon[static_cast<int64_t>(b[j]) * c.E * c.Le + k];
      for (int k = c.IN; k < c.IN4; ++k) xrow[j][k] = 0.f;
    }
    float dnext[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) dnext[j] = demand_at(c, demands, b[j], 0);
    for (int t = 0; t < c.T; ++t) {
      float d[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        d[j] = dnext[j];
        if (t + 1 < c.T) dnext[j] = demand_at(c, demands, b[j], t + 1);
      }
      if (tape && t % c.ckpt == 0) {  // checkpoint: the adjoint recomputes the c.ckpt - 1 states in between
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (valid[j]) {
            float4* dst =
                reinterpret_cast<float4*>(tape + (static_cast<int64_t>(t / c.ckpt) * c.B + b[j]) * c.tape_stride);
            for (int k4 = 0; k4 < c.IN4 / 4; ++k4) dst[k4] = reinterpret_cast<const float4*>(xrow[j])[k4];
          }
        }
      }
      float y[NS][kMaxOut];
      {
        const float* xin[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) xin[j] = xrow[j];
        // one hidden row per scenario is enough in the forward: every layer overwrites it in place
        mlp_fwd<NS>(c, Ws, xin, hrow, 0, y);
      }
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        Head hd;
        head_fwd<ARCH>(c, xrow[j], y[j], hd);
        const float r = env_fwd<ARCH>(c, xrow[j], d[j], hd, s[j]);
        cost[j] += r;
        if (t >= c.ignore) rep[j] += r;
        if (reward_tb && valid[j]) reward_tb[static_cast<int64_t>(t) * c.B + b[j]] = r;
      }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (!valid[j]) continue;
      cost_b[b[j]] = cost[j];
      if (report_b) report_b[b[j]] = rep[j];
      if (fin.store)
        for (int k = 0; k < c.L; ++k) fin.store[static_cast<int64_t>(b[j]) * c.L + k] = xrow[j][k];
      if (fin.warehouse)
        for (int k = 0; k < c.Lw; ++k) fin.warehouse[static_cast<int64_t>(b[j]) * c.Lw + k] = xrow[j][c.L + k];
      if (fin.echelon)
        for (int k = 0; k < c.E * c.Le; ++k)
          fin.echelon[static_cast<int64_t>(b[j]) * c.E * c.Le + k] = xrow[j][c.L + c.Lw + k];
    }
  }
}

// deterministic totals: one block, fixed-order tree in double
static __global__ void __launch_bounds__(1024) totals_kernel(const float* __restrict__ cost_b,
                                                      const float* __restrict__ report_b, int B,
                                                      double* __restrict__ totals) {
  __shared__ double s0[1024];
  __shared__ double s1[1024];
  // four independent partial sums per thread: the single-CTA pass over 2^20 scenarios is latency-bound (it took
  // 0.52 ms of a 42 ms step with one dependent load chain per thread); the order stays fixed, hence deterministic
  double a4[4] = {0.0, 0.0, 0.0, 0.0}, r4[4] = {0.0, 0.0, 0.0, 0.0};
  const int stride = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * stride < B; i += 4 * stride) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a4[j] += static_cast<double>(cost_b[i + j * stride]);
      if (report_b) r4[j] += static_cast<double>(report_b[i + j * stride]);
    }
  }
  for (; i < B; i += stride) {
    a4[0] += static_cast<double>(cost_b[i]);
    if (report_b) r4[0] += static_cast<double>(report_b[i]);
  }
  const double a = (a4[0] + a4[1]) + (a4[2] + a4[3]), r = (r4[0] + r4[1]) + (r4[2] + r4[3]);
  s0[threadIdx.x] = a;
  s1[threadIdx.x] = r;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if (static_cast<int>(threadIdx.x) < w) {
      s0[threadIdx.x] += s0[threadIdx.x + w];
      s1[threadIdx.x] += s1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = s0[0];
    totals[1] = s1[0];
  }
}

// ------------------------------------------------------------------------------------------------------------
// K2: reverse-time adjoint
// ------------------------------------------------------------------------------------------------------------

// lane tile of an [H x (4*KQ)] weight gradient: rows 4*ni..4*ni+3, columns ki*KQ..ki*KQ+KQ-1.
// Column pairs are accumulated with packed FFMA2 ((g_i, g_i) x (x_q, x_q+1)); an odd KQ keeps one scalar tail column.
template <int KQ>
struct WgradAcc {            // per-lane 4 x KQ tile of dW kept in registers for the whole kernel
  static constexpr int KP = KQ / 2;
  float2 pair[4][KP > 0 ? KP : 1];
  float tail[4];             // last column when KQ is odd
  float bias[4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int q = 0; q < (KP > 0 ? KP : 1); ++q) pair[i][q] = make_float2(0.f, 0.f);
      tail[i] = 0.f;
      bias[i] = 0.f;
    }
  }
  __device__ __forceinline__ float at(int i, int q) const {
    if ((KQ & 1) && q == KQ - 1) return tail[i];
    return (q & 1) ? pair[i][q >> 1].y : pair[i][q >> 1].x;
  }
};

template <int KQ>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ G, int gs, const float* __restrict__ X, int xs,
                                           int lane, WgradAcc<KQ>& a) {
  const int ni = lane >> 2, ki = lane & 3;
  constexpr int KP = KQ / 2;
  for (int cidx = 0; cidx < 32; ++cidx) {
    const float4 g = *reinterpret_cast<const float4*>(G + cidx * gs + 4 * ni);
    float xv[KQ];
    if (KQ % 4 == 0) {
#pragma unroll
      for (int q = 0; q < KQ / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(X + cidx * xs + ki * KQ + 4 * q);
        xv[4 * q + 0] = v.x;
        xv[4 * q + 1] = v.y;
        xv[4 * q + 2] = v.z;
        xv[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < KQ; ++q) xv[q] = X[cidx * xs + ki * KQ + q];
    }
    const float2 g0 = make_float2(g.x, g.x), g1 = make_float2(g.y, g.y), g2 = make_float2(g.z, g.z),
                 g3 = make_float2(g.w, g.w);
#pragma unroll
    for (int q = 0; q < KP; ++q) {
      const float2 x2 = make_float2(xv[2 * q], xv[2 * q + 1]);
      ffma2(a.pair[0][q], g0, x2);
      ffma2(a.pair[1][q], g1, x2);
      ffma2(a.pair[2][q], g2, x2);
      ffma2(a.pair[3][q], g3, x2);
    }
    if (KQ & 1) {
      a.tail[0] = fmaf(g.x, xv[KQ - 1], a.tail[0]);
      a.tail[1] = fmaf(g.y, xv[KQ - 1], a.tail[1]);
      a.tail[2] = fmaf(g.z, xv[KQ - 1], a.tail[2]);
      a.tail[3] = fmaf(g.w, xv[KQ - 1], a.tail[3]);
    }
    a.bias[0] += g.x;
    a.bias[1] += g.y;
    a.bias[2] += g.z;
    a.bias[3] += g.w;
  }
}

// g_in[k] = sum_n Wt[k][n] * gz[n]  (thread-local dgrad), k = 0..K-1; gz held as float2 pairs, packed FFMA2
__device__ __forceinline__ float dgrad_dot(const float* __restrict__ Wt_row, const float2 (&gz)[H / 2]) {
  const float4* w = reinterpret_cast<const float4*>(Wt_row);
  float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
  for (int n4 = 0; n4 < H / 4; ++n4) {
    const float4 wv = w[n4];
    ffma2(s01, make_float2(wv.x, wv.y), gz[2 * n4 + 0]);
    ffma2(s23, make_float2(wv.z, wv.w), gz[2 * n4 + 1]);
  }
  return (s01.x + s01.y) + (s23.x + s23.y);
}

template <int ARCH, int KQ0, int NHH, bool TC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
small_bwd_kernel(Cfg c, const float* __restrict__ params, const float* __restrict__ demands, HdpoStatics st,
                 const float* __restrict__ tape, float g_total, float g_report, float* __restrict__ partials,
                 int p_stride) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights(c, params, Ws, true);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NHB = NHH + 1;  // hidden rows kept per scenario
  const int per_warp = 32 * (2 * c.XSb + NHB * HS + kMaxOut);
  float* Xw = smem + ((c.s_total_bwd + 3) & ~3) + warp * per_warp;  // state rows x_t
  float* Gx = Xw + 32 * c.XSb;                                   // state adjoint rows
  float* Hb = Gx + 32 * c.XSb;                                   // [NHB][32][HS] activations -> overwritten by gz
  float* Gy = Hb + NHB * 32 * HS;                               // [32][kMaxOut] output adjoints
  float* xrow = Xw + lane * c.XSb;
  float* grow = Gx + lane * c.XSb;
  float* hrow = Hb + lane * HS;
  constexpr int HL = 32 * HS;  // layer stride inside Hb

  // register-resident parameter-gradient tiles
  WgradAcc<TC ? 1 : KQ0> a0;
  WgradAcc<TC ? 1 : 8> ah[NHH > 0 ? NHH : 1];
  constexpr int NT0 = (KQ0 + 1) / 2;  // 8-column blocks of the zero-padded first-layer inputs (K0 = 8 NT0)
  float a0f[2][TC ? NT0 : 1][4], b0f = 0.f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < (TC ? NT0 : 1); ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) a0f[mt][nt][i] = 0.f;
  float ao[kMaxOut], bo = 0.f;  // lane = k for ao; lane = o for bo
  a0.clear();
#pragma unroll
  for (int l = 0; l < (NHH > 0 ? NHH : 1); ++l) ah[l].clear();
  // tensor-core mode: the HxH weight gradients are mma C fragments (element i of [mt][nt]: n = 16 mt + (lane >> 2) +
  // 8 (i >> 1), k = 8 nt + 2 (lane & 3) + (i & 1)), their bias sums live with lane = n
  float ahf[NHH > 0 ? NHH : 1][2][TC ? 4 : 1][4], bhf[NHH > 0 ? NHH : 1];
#pragma unroll
  for (int l = 0; l < (NHH > 0 ? NHH : 1); ++l) {
    bhf[l] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < (TC ? 4 : 1); ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) ahf[l][mt][nt][i] = 0.f;
  }
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) ao[o] = 0.f;

  const int n_tiles = ceil_div(c.B, 32);
  const int wpc = blockDim.x >> 5;  // warps per CTA: 4 for big batches, 2 / 1 when there are few tiles (latency)
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  for (int k = c.IN4; k < c.XSb; ++k) xrow[k] = 0.f;  // padding columns (read by the mma first layer) stay zero
  for (int tile = gwarp; tile < n_tiles; tile += nwarps) {
    const int bb = tile * 32 + lane;
    const bool valid = bb < c.B;
    const int b = valid ? bb : c.B - 1;
    Statics s;
    load_statics<ARCH>(c, st, b, s);
    for (int k = 0; k < c.XSb; ++k) grow[k] = 0.f;
    // Recomputation checkpointing (c.ckpt = K > 1): the tape holds x_0, x_K, x_2K, ...; a segment [seg0, seg1) is
    // re-run forward from its checkpoint (same device functions as the forward kernel: bit-identical states), the
    // states x_{seg0+1 .. seg1-1} go to this warp's ring (a few KB per warp, L2-resident), then the reverse sweep of
    // the segment reads them back. K = 1 degenerates to the plain tape.
    float* ring = c.ring ? c.ring + static_cast<int64_t>(gwarp) * c.ckpt * 32 * c.tape_stride : nullptr;
    for (int seg0 = ((c.T - 1) / c.ckpt) * c.ckpt; seg0 >= 0; seg0 -= c.ckpt) {
    const int seg1 = seg0 + c.ckpt < c.T ? seg0 + c.ckpt : c.T;
    if (seg1 - seg0 > 1) {
      const float4* src =
          reinterpret_cast<const float4*>(tape + (static_cast<int64_t>(seg0 / c.ckpt) * c.B + b) * c.tape_stride);
      for (int k4 = 0; k4 < c.IN4 / 4; ++k4) reinterpret_cast<float4*>(xrow)[k4] = src[k4];
      for (int t = seg0; t + 1 < seg1; ++t) {
        float yf[1][kMaxOut];
        const float* xin[1] = {xrow};
        float* hr[1] = {hrow};
        __syncwarp();
        mlp_fwd<1>(c, Ws, xin, hr, HL, yf);
        Head hf;
        head_fwd<ARCH>(c, xrow, yf[0], hf);
        env_fwd<ARCH>(c, xrow, demand_at(c, demands, b, t), hf, s);
        float4* dst = reinterpret_cast<float4*>(ring + (static_cast<int64_t>(t + 1 - seg0) * 32 + lane) * c.tape_stride);
        for (int k4 = 0; k4 < c.IN4 / 4; ++k4) dst[k4] = reinterpret_cast<const float4*>(xrow)[k4];
      }
      __syncwarp();
    }
    for (int t = seg1 - 1; t >= seg0; --t) {
      // A. state x_t from the tape (checkpoint) or from the segment ring, demand
      {
        const float4* src =
            t == seg0 ? reinterpret_cast<const float4*>(tape + (static_cast<int64_t>(seg0 / c.ckpt) * c.B + b) * c.tape_stride)
                      : reinterpret_cast<const float4*>(ring + (static_cast<int64_t>(t - seg0) * 32 + lane) * c.tape_stride);
        for (int k4 = 0; k4 < c.IN4 / 4; ++k4) reinterpret_cast<float4*>(xrow)[k4] = src[k4];
      }
      const float d = demand_at(c, demands, b, t);
      const float rb = valid ? (g_total + (t >= c.ignore ? g_report : 0.f)) : 0.f;
      // B. recompute the activations (rows hb[0..NHH]) and the outputs
      float y[1][kMaxOut];
      if constexpr (TC) {
#ifndef HDPO_EMU
        mlp_fwd_tc(c, Ws, Xw, Hb, HL, lane, y[0]);
#endif
      } else {
        const float* xin[1] = {xrow};
        float* hr[1] = {hrow};
        mlp_fwd<1>(c, Ws, xin, hr, HL, y);
      }
      // C. head + simulator adjoint (thread local); padded lanes carry rb = 0 and a zero adjoint row
      Head hd;
      head_fwd<ARCH>(c, xrow, y[0], hd);
      float gy[kMaxOut];
      head_env_bwd<ARCH>(c, xrow, grow, d, hd, s, rb, gy);
      if (!valid) {
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) gy[o] = 0.f;
      }
#pragma unroll
      for (int o4 = 0; o4 < kMaxOut / 4; ++o4)
        reinterpret_cast<float4*>(Gy + lane * kMaxOut)[o4] =
            make_float4(gy[4 * o4], gy[4 * o4 + 1], gy[4 * o4 + 2], gy[4 * o4 + 3]);
      __syncwarp();
      // D. output layer: dWo[o][k] += sum_c gy[c][o] * h_last[c][k]  (lane = k), dbo[o] (lane = o)
      {
        const float* Hl = Hb + NHH * HL;
        for (int cidx = 0; cidx < 32; ++cidx) {
          const float hv = Hl[cidx * HS + lane];
          const float4 g0v = *reinterpret_cast<const float4*>(Gy + cidx * kMaxOut);
          const float4 g1v = *reinterpret_cast<const float4*>(Gy + cidx * kMaxOut + 4);
          ao[0] = fmaf(g0v.x, hv, ao[0]);
          ao[1] = fmaf(g0v.y, hv, ao[1]);
          ao[2] = fmaf(g0v.z, hv, ao[2]);
          ao[3] = fmaf(g0v.w, hv, ao[3]);
          ao[4] = fmaf(g1v.x, hv, ao[4]);
          ao[5] = fmaf(g1v.y, hv, ao[5]);
          ao[6] = fmaf(g1v.z, hv, ao[6]);
          ao[7] = fmaf(g1v.w, hv, ao[7]);
          if (lane < kMaxOut) bo += Gy[cidx * kMaxOut + lane];
        }
      }
      __syncwarp();
      // thread-local: g_h_last[k] = sum_o Wo[o][k] gy[o]; gz = g_h * act'(h) -> overwrite the activation row
      float2 gz[H / 2];
      {
        float* hl = hrow + NHH * HL;
        dispatch_act(c.hidden_act, [&](auto tag) {
          constexpr int ACT = decltype(tag)::value;
#pragma unroll
          for (int k4 = 0; k4 < H / 4; ++k4) {
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < kMaxOut; ++o) {
              if (o < c.OUT) {
                const float4 wv = reinterpret_cast<const float4*>(Ws + c.s_wo + o * H)[k4];
                acc4.x = fmaf(wv.x, gy[o], acc4.x);
                acc4.y = fmaf(wv.y, gy[o], acc4.y);
                acc4.z = fmaf(wv.z, gy[o], acc4.z);
                acc4.w = fmaf(wv.w, gy[o], acc4.w);
              }
            }
            const float4 hv = reinterpret_cast<const float4*>(hl)[k4];
            gz[2 * k4 + 0] = make_float2(acc4.x * act_grad_out_t<ACT>(hv.x), acc4.y * act_grad_out_t<ACT>(hv.y));
            gz[2 * k4 + 1] = make_float2(acc4.z * act_grad_out_t<ACT>(hv.z), acc4.w * act_grad_out_t<ACT>(hv.w));
            reinterpret_cast<float4*>(hl)[k4] =
                make_float4(gz[2 * k4].x, gz[2 * k4].y, gz[2 * k4 + 1].x, gz[2 * k4 + 1].y);
          }
        });
      }
      // E. hidden HxH layers, top down
#pragma unroll
      for (int l = NHH - 1; l >= 0; --l) {
        __syncwarp();
        if constexpr (TC) {
#ifndef HDPO_EMU
          const float* Gz = Hb + (l + 1) * HL;
          mma32::wgrad<4>(Gz, HS, Hb + l * HL, HS, lane, ahf[l]);
          float bs = 0.f;
          for (int cidx = 0; cidx < 32; ++cidx) bs += Gz[cidx * HS + lane];
          bhf[l] += bs;
          dispatch_act(c.hidden_act, [&](auto tag) {
            constexpr int ACT = decltype(tag)::value;
            mma32::dgrad_inplace_f32(Ws + c.s_wn[l], HS, Gz, HS, Hb + l * HL, HS, lane,
                                     [](float yv) { return act_grad_out_t<ACT>(yv); });
          });
          float* hq = hrow + l * HL;
#pragma unroll
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 v = reinterpret_cast<const float4*>(hq)[k4];
            gz[2 * k4 + 0] = make_float2(v.x, v.y);
            gz[2 * k4 + 1] = make_float2(v.z, v.w);
          }
#endif
        } else {
        wgrad_tile<8>(Hb + (l + 1) * HL, HS, Hb + l * HL, HS, lane, ah[l]);
        __syncwarp();
        float* hp = hrow + l * HL;
        const float* Wt = Ws + c.s_wth[l];
        // thread-local dgrad: gz_{l-1}[k] = (sum_n Wt[k][n] gz_l[n]) * act'(h_{l-1}[k]), written in place over h_{l-1}
        dispatch_act(c.hidden_act, [&](auto tag) {
          constexpr int ACT = decltype(tag)::value;
#pragma unroll 1
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 hv = reinterpret_cast<const float4*>(hp)[k4];
            float4 r;
            r.x = dgrad_dot(Wt + (4 * k4 + 0) * H, gz) * act_grad_out_t<ACT>(hv.x);
            r.y = dgrad_dot(Wt + (4 * k4 + 1) * H, gz) * act_grad_out_t<ACT>(hv.y);
            r.z = dgrad_dot(Wt + (4 * k4 + 2) * H, gz) * act_grad_out_t<ACT>(hv.z);
            r.w = dgrad_dot(Wt + (4 * k4 + 3) * H, gz) * act_grad_out_t<ACT>(hv.w);
            reinterpret_cast<float4*>(hp)[k4] = r;
          }
        });
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 v = reinterpret_cast<const float4*>(hp)[k4];
          gz[2 * k4 + 0] = make_float2(v.x, v.y);
          gz[2 * k4 + 1] = make_float2(v.z, v.w);
        }
        }  // FFMA form
      }
      // F. layer 0: dW0 += gz0^T x ; state adjoint += W0^T gz0 unless the input was detached
      __syncwarp();
      if constexpr (TC) {
#ifndef HDPO_EMU
        mma32::wgrad<NT0>(Hb, HS, Xw, c.XSb, lane, a0f);
        float bs = 0.f;
        for (int cidx = 0; cidx < 32; ++cidx) bs += Hb[cidx * HS + lane];
        b0f += bs;
        if (!c.detach_input) mma32::dgrad_accum_f32<NT0>(Ws + c.s_w0n, c.XSb, Hb, HS, Gx, c.XSb, lane);
#endif
      } else {
        wgrad_tile<KQ0>(Hb, HS, Xw, c.XSb, lane, a0);
        __syncwarp();
        if (!c.detach_input) {
          const float* Wt0 = Ws + c.s_wt0;
          for (int k = 0; k < c.IN; ++k) grow[k] += dgrad_dot(Wt0 + k * H, gz);
        }
      }
    }
    }  // segments
  }

  // ---- write this warp's partial gradient slab in state_dict layout
  float* out = partials + static_cast<int64_t>(gwarp) * p_stride;
  const int ni = lane >> 2, ki = lane & 3;
  if constexpr (TC) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < (TC ? NT0 : 1); ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = 16 * mt + ni + 8 * (i >> 1), k = 8 * nt + 2 * ki + (i & 1);
          if (n < c.w[1] && k < c.IN) out[c.gw[0] + n * c.IN + k] = a0f[mt][nt][i];
        }
    if (lane < c.w[1]) out[c.gb[0] + lane] = b0f;
  } else {  // layer 0: W0[n][k], n = 4*ni+i, k = ki*KQ0+q
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = 4 * ni + i;
      if (n < c.w[1]) {
#pragma unroll
        for (int q = 0; q < (TC ? 1 : KQ0); ++q) {
          const int k = ki * KQ0 + q;
          if (k < c.IN) out[c.gw[0] + n * c.IN + k] = a0.at(i, q);
        }
        if (ki == 0) out[c.gb[0] + n] = a0.bias[i];
      }
    }
  }
#pragma unroll
  for (int l = 0; l < NHH; ++l) {
    const int n_out = c.w[l + 2], n_in = c.w[l + 1];
    if constexpr (TC) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < (TC ? 4 : 1); ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = 16 * mt + ni + 8 * (i >> 1), k = 8 * nt + 2 * ki + (i & 1);
            if (n < n_out && k < n_in) out[c.gw[l + 1] + n * n_in + k] = ahf[l][mt][nt][i];
          }
      if (lane < n_out) out[c.gb[l + 1] + lane] = bhf[l];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = 4 * ni + i;
        if (n < n_out) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = ki * 8 + q;
            if (k < n_in) out[c.gw[l + 1] + n * n_in + k] = ah[l].at(i, q);
          }
          if (ki == 0) out[c.gb[l + 1] + n] = ah[l].bias[i];
        }
      }
    }
  }
  {
    const int n_in = c.w[NHH + 1];
#pragma unroll
    for (int o = 0; o < kMaxOut; ++o)
      if (o < c.OUT && lane < n_in) out[c.gw[NHH + 1] + o * n_in + lane] = ao[o];
    if (lane < c.OUT) out[c.gb[NHH + 1] + lane] = bo;
  }
}

// grad[p] = sum over slabs in fixed order (deterministic)
static __global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n_rows,
                                                              int p_stride, int P, float* __restrict__ grad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = 0;
  for (; r + 3 < n_rows; r += 4) {
    s0 += partials[static_cast<int64_t>(r) * p_stride + p];
    s1 += partials[static_cast<int64_t>(r + 1) * p_stride + p];
    s2 += partials[static_cast<int64_t>(r + 2) * p_stride + p];
    s3 += partials[static_cast<int64_t>(r + 3) * p_stride + p];
  }
  for (; r < n_rows; ++r) s0 += partials[static_cast<int64_t>(r) * p_stride + p];
  grad[p] = (s0 + s1) + (s2 + s3);
}


// launch of one adjoint instantiation (defined per <ARCH, KQ0> in the rollout_small_bwd_*.cu units)
template <int ARCH, int KQ0>
int launch_bwd_nhh(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
                   float g_total, float g_report, float* partials, int p_stride, int grid, int wpc, void* stream);

template <int ARCH, int KQ0, int NHH, bool TC>
int launch_bwd_tc(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
                  float g_total, float g_report, float* partials, int p_stride, int grid, int wpc, void* stream) {
  const int per_warp = 32 * (2 * c.XSb + (NHH + 1) * HS + kMaxOut);
  const size_t wfl = static_cast<size_t>((c.s_total_bwd + 3) & ~3);
  const size_t smem = (wfl + static_cast<size_t>(wpc) * per_warp) * sizeof(float);
  auto k = small_bwd_kernel<ARCH, KQ0, NHH, TC>;
  HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>((wfl + static_cast<size_t>(kWarpsPerCta) * per_warp) * sizeof(float))));
  HDPO_LAUNCH(k, grid, wpc * 32, smem, stream, c, params, demands, *st, tape, g_total, g_report, partials, p_stride);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}
template <int ARCH, int KQ0, int NHH>
int launch_bwd(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
               float g_total, float g_report, float* partials, int p_stride, int grid, int wpc, void* stream) {
#ifndef HDPO_EMU
  if (NHH > 0 && c.tc)
    return launch_bwd_tc<ARCH, KQ0, NHH, (NHH > 0)>(c, params, demands, st, tape, g_total, g_report, partials, p_stride,
                                                    grid, wpc, stream);
#endif
  return launch_bwd_tc<ARCH, KQ0, NHH, false>(c, params, demands, st, tape, g_total, g_report, partials, p_stride, grid,
                                              wpc, stream);
}

#define HDPO_SMALL_BWD_INSTANCE(ARCH, KQ0)                                                                         \
  template <>                                                                                                      \
  int launch_bwd_nhh<ARCH, KQ0>(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st,    \
                                const float* tape, float g_total, float g_report, float* partials, int p_stride,   \
                                int grid, int wpc, void* stream) {                                                 \
    switch (c.NHH) {                                                                                               \
      case 0: return launch_bwd<ARCH, KQ0, 0>(c, params, demands, st, tape, g_total, g_report, partials, p_stride, grid, wpc, stream); \
      case 1: return launch_bwd<ARCH, KQ0, 1>(c, params, demands, st, tape, g_total, g_report, partials, p_stride, grid, wpc, stream); \
      case 2: return launch_bwd<ARCH, KQ0, 2>(c, params, demands, st, tape, g_total, g_report, partials, p_stride, grid, wpc, stream); \
      case 3: return launch_bwd<ARCH, KQ0, 3>(c, params, demands, st, tape, g_total, g_report, partials, p_stride, grid, wpc, stream); \
    }                                                                                                              \
    set_error("unsupported hidden depth %d", c.NHH);                                                               \
    return HDPO_E_INVALID;                                                                                         \
  }

}  // namespace small
}  // namespace hdpo
