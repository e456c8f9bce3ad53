// rollout_small_unit.cu - the small-net rollout (K1/K2, one-store and serial policies) for TRAINING-SIZE batches.
//
// rollout_small_kernels.cuh maps lane = scenario: a warp owns 32 scenarios and walks their T periods, which is the
// cheapest form per scenario (weights broadcast from shared memory, 32 independent accumulators per lane) but leaves a
// batch of 8192 scenarios with 256 warps of work - fewer than two per SM, each a 50-period dependent chain. Here the
// mapping is lane = HIDDEN UNIT and ONE SCENARIO PER WARP: a 32 x 32 layer is 32 FMAs per lane against the lane's
// weight row (W[n][k], row stride 36: conflict-free LDS.128) and the broadcast activation vector, so 8192 scenarios are
// B independent warp tasks; it is used while the lane = scenario form cannot fill the machine (unit_max_batch below). The policy head + simulator period (a few dozen scalar operations on the scenario's state row) run on
// lane 0 with the device functions of the lane = scenario kernels, so both forms produce the same states and tapes;
// the adjoint keeps the weight gradients of its warp in shared memory (row stride 36, lane = output unit) and writes
// one slab per warp at the end, reduced in a fixed order like the other form (deterministic gradient).
// Reference semantics: trainer.py:181-216, environment.py:110-299, neural_networks.py:200-214 / 319-355.
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace small {

constexpr int kUnitWarpsFwd = 16;
constexpr int kUnitWarpsBwd = 12;
// dispatch thresholds (scenarios): largest batch of the lane = unit kernels per policy
constexpr int kUnitMaxOneStore = 4096;
constexpr int kUnitMaxSerial = 2048;

// shared-memory weight block of the unit kernels (float offsets). "n" = W[n][k] rows (forward, weight gradient),
// "t" = W^T[k][n] rows (input gradient); all rows zero padded to 32 entries, stride HS (first layer: s0)
struct UnitLayout {
  int s0, w0n, w0t, b0;
  int whn[kMaxHH], wht[kMaxHH], bh[kMaxHH];
  int wo, bo;
  int total;
  // per-warp gradient accumulators of the adjoint (same row strides)
  int a_w0, a_b0, a_wh[kMaxHH], a_bh[kMaxHH], a_wo, a_bo, a_total;
};

static UnitLayout unit_layout(const Cfg& c) {
  UnitLayout u{};
  int o = 0;
  auto take = [&](int n) {
    const int at = o;
    o += (n + 3) & ~3;
    return at;
  };
  u.s0 = c.IN4 | 4;  // row stride with an odd number of 16-byte chunks: conflict-free LDS.128 for lane = row
  u.w0n = take(H * u.s0);
  u.w0t = take(kMaxIn * HS);
  u.b0 = take(H);
  for (int l = 0; l < c.NHH; ++l) {
    u.whn[l] = take(H * HS);
    u.wht[l] = take(H * HS);
    u.bh[l] = take(H);
  }
  u.wo = take(kMaxOut * HS);
  u.bo = take(kMaxOut);
  u.total = o;
  o = 0;
  u.a_w0 = take(H * u.s0);
  u.a_b0 = take(H);
  for (int l = 0; l < c.NHH; ++l) {
    u.a_wh[l] = take(H * HS);
    u.a_bh[l] = take(H);
  }
  u.a_wo = take(kMaxOut * HS);
  u.a_bo = take(kMaxOut);
  u.a_total = o;
  return u;
}

static __device__ void stage_weights_unit(const Cfg& c, const UnitLayout& u, const float* __restrict__ params,
                                          float* __restrict__ Ws) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < u.total; i += nt) Ws[i] = 0.f;
  __syncthreads();
  {
    const int n_out = c.w[1], n_in = c.IN;
    for (int i = tid; i < n_out * n_in; i += nt) {
      const int n = i / n_in, k = i % n_in;
      const float w = params[c.gw[0] + i];
      Ws[u.w0n + n * u.s0 + k] = w;
      Ws[u.w0t + k * HS + n] = w;
    }
    for (int i = tid; i < n_out; i += nt) Ws[u.b0 + i] = params[c.gb[0] + i];
  }
  for (int l = 0; l < c.NHH; ++l) {
    const int n_out = c.w[l + 2], n_in = c.w[l + 1];
    for (int i = tid; i < n_out * n_in; i += nt) {
      const int n = i / n_in, k = i % n_in;
      const float w = params[c.gw[l + 1] + i];
      Ws[u.whn[l] + n * HS + k] = w;
      Ws[u.wht[l] + k * HS + n] = w;
    }
    for (int i = tid; i < n_out; i += nt) Ws[u.bh[l] + i] = params[c.gb[l + 1] + i];
  }
  {
    const int n_out = c.OUT, n_in = c.w[c.NHH + 1];
    for (int i = tid; i < n_out * n_in; i += nt) {
      const int o = i / n_in, k = i % n_in;
      Ws[u.wo + o * HS + k] = params[c.gw[c.NHH + 1] + i];
    }
    for (int i = tid; i < n_out; i += nt) Ws[u.bo + i] = params[c.gb[c.NHH + 1] + i];
  }
  __syncthreads();
}

// G scenarios per warp (G = 1, 2, 4): the lane's weight row is loaded ONCE per float4 chunk and used for the G activation
// vectors of the warp's scenarios (G independent FMA chains per lane), the G policy heads + simulator periods run on
// lanes 0 .. G-1 at the same time, and the warp's weight-gradient rows are read-modified-written once for G scenarios.
// G = 1 is the one-scenario-per-warp form and the DEFAULT at every batch size: G = 2 / 4 execute fewer warp instructions
// per scenario (the lane-0 head section was a third of them) but the period is a chain of dependent phases whose latency
// grows with G, and at 12 - 16 warps per SM that is what bounds the kernels. Measured on B200 (fwd + adjoint of 50
// periods, ms per step, G = 1 / 2 / 4 | lane = scenario form): one-store 1024: 0.53 / 0.68 / 1.10 | 1.06; 4096: 0.92 / 0.91 /
// 1.12 | 1.11; 8192: 1.53 / 1.30 / 1.45 | 1.12; 32768: 5.3 / 4.1 / 3.5 | 1.76; serial 8192: 2.04 / 1.48 / 1.45 | 1.05. G > 1 stays
// available for A/B runs (HDPO_SMALL_UNIT_G, hdpo_debug_set_small_unit_group) and is parity-tested like G = 1.
constexpr int XR = kMaxIn + 4;          // floats of one scenario's state row (and of its adjoint row)
constexpr int HSZ = (kMaxHH + 1) * H;   // floats of one scenario's activation vectors

// out[g] = z + sum_k row[k] * vec_g[k] over K4 float4 chunks: this lane's own row against G vectors every lane reads
template <int G>
__device__ __forceinline__ void row_dot_g(const float* __restrict__ row, const float* __restrict__ vec, int vstride,
                                          int K4, float z, float (&out)[G]) {
  float s0[G], s1[G], s2[G], s3[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    s0[g] = z;
    s1[g] = s2[g] = s3[g] = 0.f;
  }
#pragma unroll 8
  for (int k4 = 0; k4 < K4; ++k4) {
    const float4 w = reinterpret_cast<const float4*>(row)[k4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float4 v = reinterpret_cast<const float4*>(vec + g * vstride)[k4];
      s0[g] = fmaf(w.x, v.x, s0[g]);
      s1[g] = fmaf(w.y, v.y, s1[g]);
      s2[g] = fmaf(w.z, v.z, s2[g]);
      s3[g] = fmaf(w.w, v.w, s3[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) out[g] = (s0[g] + s1[g]) + (s2[g] + s3[g]);
}
// row[k] += sum_g gz[g] * vec_g[k] over K4 float4 chunks (this lane's accumulator row; scenarios added in order g)
template <int G>
__device__ __forceinline__ void row_axpy_g(float* __restrict__ row, const float* __restrict__ vec, int vstride, int K4,
                                           const float (&gz)[G]) {
#pragma unroll 8
  for (int k4 = 0; k4 < K4; ++k4) {
    float4 a = reinterpret_cast<float4*>(row)[k4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float4 v = reinterpret_cast<const float4*>(vec + g * vstride)[k4];
      a.x = fmaf(gz[g], v.x, a.x);
      a.y = fmaf(gz[g], v.y, a.y);
      a.z = fmaf(gz[g], v.z, a.z);
      a.w = fmaf(gz[g], v.w, a.w);
    }
    reinterpret_cast<float4*>(row)[k4] = a;
  }
}

// MLP forward of the warp's G scenarios: x = G state rows (stride XR), hs = G x (NHH + 1) activation vectors (stride
// HSZ per scenario), hreg[g][l] = this lane's unit of hidden layer l of scenario g; the outputs go to ys[g][kMaxOut]
// (shared memory: the lane that runs scenario g's head reads them after the final __syncwarp()).
template <int G>
__device__ __forceinline__ void unit_mlp_fwd(const Cfg& c, const UnitLayout& u, const float* __restrict__ Ws,
                                             const float* __restrict__ x, float* __restrict__ hs,
                                             float* __restrict__ ys, int lane, float (&hreg)[G][kMaxHH + 1]) {
  float z[G];
  row_dot_g<G>(Ws + u.w0n + lane * u.s0, x, XR, c.IN4 / 4, Ws[u.b0 + lane], z);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float h = act_fwd(c.hidden_act, z[g]);
    hreg[g][0] = h;
    hs[g * HSZ + lane] = h;
  }
  __syncwarp();
#pragma unroll
  for (int l = 0; l < kMaxHH; ++l) {
    if (l < c.NHH) {
      row_dot_g<G>(Ws + u.whn[l] + lane * HS, hs + l * H, HSZ, H / 4, Ws[u.bh[l] + lane], z);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float h = act_fwd(c.hidden_act, z[g]);
        hreg[g][l + 1] = h;
        hs[g * HSZ + (l + 1) * H + lane] = h;
      }
      __syncwarp();
    }
  }
  const int o = lane < c.OUT ? lane : 0;
  row_dot_g<G>(Ws + u.wo + o * HS, hs + c.NHH * H, HSZ, H / 4, Ws[u.bo + o], z);
  if (lane < kMaxOut) {
#pragma unroll
    for (int g = 0; g < G; ++g) ys[g * kMaxOut + lane] = lane < c.OUT ? z[g] : 0.f;
  }
  __syncwarp();
}

template <int ARCH, int G>
__global__ void __launch_bounds__(kUnitWarpsFwd * 32, 1)
small_unit_fwd_kernel(Cfg c, UnitLayout u, const float* __restrict__ params, const float* __restrict__ demands,
                      HdpoStatics st, HdpoState init, float* __restrict__ cost_b, float* __restrict__ report_b,
                      float* __restrict__ reward_tb, float* __restrict__ tape, HdpoState fin) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights_unit(c, u, params, Ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  constexpr int per_warp = G * (XR + HSZ + kMaxOut);
  float* x = smem + u.total + warp * per_warp;  // [G][XR]
  float* hs = x + G * XR;                       // [G][HSZ]
  float* ys = hs + G * HSZ;                     // [G][kMaxOut]
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  const int n_tiles = (c.B + G - 1) / G;
  const int gl = lane < G ? lane : 0;  // the scenario of the tile whose head / simulator period this lane runs
  float* xme = x + gl * XR;
  for (int tile = gwarp; tile < n_tiles; tile += nwarps) {
    const int b0 = tile * G;
    const bool active = lane < G && b0 + lane < c.B;
    const int bme = b0 + gl < c.B ? b0 + gl : c.B - 1;
    Statics s;
    load_statics<ARCH>(c, st, bme, s);
    __syncwarp();
    // initial state rows: [store L | warehouse Lw | echelons E*Le], zero padded to IN4 (tail scenarios: all zero)
    for (int i = lane; i < G * c.IN4; i += 32) {
      const int g = i / c.IN4, k = i - g * c.IN4;
      const int64_t b = b0 + g;
      float v = 0.f;
      if (b < c.B) {
        if (k < c.L) v = init.store[b * c.L + k];
        else if (k < c.L + c.Lw) v = init.warehouse[b * c.Lw + (k - c.L)];
        else if (k < c.IN) v = init.echelon[b * c.E * c.Le + (k - c.L - c.Lw)];
      }
      x[g * XR + k] = v;
    }
    __syncwarp();
    float cost = 0.f, rep = 0.f;
    float dnext = demand_at(c, demands, bme, 0);
    const int n4 = c.IN4 / 4;
    for (int t = 0; t < c.T; ++t) {
      const float d = dnext;
      if (t + 1 < c.T) dnext = demand_at(c, demands, bme, t + 1);
      if (tape && t % c.ckpt == 0) {
        for (int i = lane; i < G * n4; i += 32) {
          const int g = i / n4, j = i - g * n4;
          if (b0 + g < c.B)
            reinterpret_cast<float4*>(tape + (static_cast<int64_t>(t / c.ckpt) * c.B + b0 + g) * c.tape_stride)[j] =
                reinterpret_cast<const float4*>(x + g * XR)[j];
        }
      }
      float hreg[G][kMaxHH + 1];
      unit_mlp_fwd<G>(c, u, Ws, x, hs, ys, lane, hreg);
      if (active) {
        float y[kMaxOut];
#pragma unroll
        for (int i = 0; i < kMaxOut; ++i) y[i] = ys[gl * kMaxOut + i];
        Head hd;
        head_fwd<ARCH>(c, xme, y, hd);
        const float r = env_fwd<ARCH>(c, xme, d, hd, s);
        cost += r;
        if (t >= c.ignore) rep += r;
        if (reward_tb) reward_tb[static_cast<int64_t>(t) * c.B + b0 + lane] = r;
      }
      __syncwarp();
    }
    if (active) {
      cost_b[b0 + lane] = cost;
      if (report_b) report_b[b0 + lane] = rep;
    }
    for (int i = lane; i < G * c.IN; i += 32) {
      const int g = i / c.IN, k = i - g * c.IN;
      const int64_t b = b0 + g;
      if (b >= c.B) continue;
      const float v = x[g * XR + k];
      if (k < c.L) {
        if (fin.store) fin.store[b * c.L + k] = v;
      } else if (k < c.L + c.Lw) {
        if (fin.warehouse) fin.warehouse[b * c.Lw + (k - c.L)] = v;
      } else if (fin.echelon) {
        fin.echelon[b * c.E * c.Le + (k - c.L - c.Lw)] = v;
      }
    }
    __syncwarp();
  }
}

template <int ARCH, int G>
__global__ void __launch_bounds__(kUnitWarpsBwd * 32, 1)
small_unit_bwd_kernel(Cfg c, UnitLayout u, const float* __restrict__ params, const float* __restrict__ demands,
                      HdpoStatics st, const float* __restrict__ tape, float g_total, float g_report,
                      float* __restrict__ partials, int p_stride) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights_unit(c, u, params, Ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int per_warp = u.a_total + G * (2 * XR + HSZ + H + kMaxOut);
  float* acc = smem + u.total + warp * per_warp;
  float* x = acc + u.a_total;   // [G][XR] state rows
  float* gr = x + G * XR;       // [G][XR] state adjoint rows
  float* hs = gr + G * XR;      // [G][HSZ]
  float* gzs = hs + G * HSZ;    // [G][H]
  float* gys = gzs + G * H;     // [G][kMaxOut] (also the outputs ys of the recomputed forward)
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  for (int i = lane; i < u.a_total; i += 32) acc[i] = 0.f;
  __syncwarp();
  const int n_tiles = (c.B + G - 1) / G;
  const int gl = lane < G ? lane : 0;
  const int n4 = c.IN4 / 4;
  for (int tile = gwarp; tile < n_tiles; tile += nwarps) {
    const int b0 = tile * G;
    const bool active = lane < G && b0 + lane < c.B;
    const int bme = b0 + gl < c.B ? b0 + gl : c.B - 1;
    Statics s;
    load_statics<ARCH>(c, st, bme, s);
    __syncwarp();
    for (int i = lane; i < G * XR; i += 32) gr[i] = 0.f;  // adjoint wrt the state after the last period
    for (int t = c.T - 1; t >= 0; --t) {
      for (int i = lane; i < G * n4; i += 32) {
        const int g = i / n4, j = i - g * n4;
        const int b = b0 + g < c.B ? b0 + g : c.B - 1;  // (tail scenarios recompute a valid row; their adjoint is zero)
        reinterpret_cast<float4*>(x + g * XR)[j] =
            reinterpret_cast<const float4*>(tape + (static_cast<int64_t>(t) * c.B + b) * c.tape_stride)[j];
      }
      const float d = demand_at(c, demands, bme, t);
      __syncwarp();
      float hreg[G][kMaxHH + 1];
      unit_mlp_fwd<G>(c, u, Ws, x, hs, gys, lane, hreg);
      if (lane < G) {
        float gy[kMaxOut];
        if (active) {
          float y[kMaxOut];
#pragma unroll
          for (int i = 0; i < kMaxOut; ++i) y[i] = gys[gl * kMaxOut + i];
          Head hd;
          head_fwd<ARCH>(c, x + gl * XR, y, hd);
          const float rb = g_total + (t >= c.ignore ? g_report : 0.f);
          head_env_bwd<ARCH>(c, x + gl * XR, gr + gl * XR, d, hd, s, rb, gy);
        } else {
#pragma unroll
          for (int i = 0; i < kMaxOut; ++i) gy[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < kMaxOut; ++i) gys[lane * kMaxOut + i] = gy[i];
      }
      __syncwarp();
      // ---- output layer: dWo[o][k] += gy[o] h[k] (lane = k), dbo, gh[k] = sum_o Wo[o][k] gy[o]
      float gh[G];
      {
        float hl[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          gh[g] = 0.f;
          hl[g] = hreg[g][0];
#pragma unroll
          for (int l = 1; l <= kMaxHH; ++l)
            if (l == c.NHH) hl[g] = hreg[g][l];
        }
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          if (o < c.OUT) {
            float a = acc[u.a_wo + o * HS + lane];
            const float w = Ws[u.wo + o * HS + lane];
            float bsum = 0.f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const float gyo = gys[g * kMaxOut + o];
              a = fmaf(gyo, hl[g], a);
              gh[g] = fmaf(w, gyo, gh[g]);
              bsum += gyo;
            }
            acc[u.a_wo + o * HS + lane] = a;
            if (lane == o) acc[u.a_bo + o] += bsum;
          }
        }
      }
      // ---- hidden layers, last to first: gz = gh * act'(h); dW[n][:] += gz[n] h_prev[:]; gh_prev = W^T gz
#pragma unroll
      for (int l = kMaxHH - 1; l >= 0; --l) {
        if (l < c.NHH) {
          float gz[G], bsum = 0.f;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            gz[g] = gh[g] * act_grad_from_out(c.hidden_act, hreg[g][l + 1]);
            gzs[g * H + lane] = gz[g];
            bsum += gz[g];
          }
          acc[u.a_bh[l] + lane] += bsum;
          row_axpy_g<G>(acc + u.a_wh[l] + lane * HS, hs + l * H, HSZ, H / 4, gz);
          __syncwarp();
          row_dot_g<G>(Ws + u.wht[l] + lane * HS, gzs, H, H / 4, 0.f, gh);
          __syncwarp();
        }
      }
      // ---- first layer: weight gradient against the state rows, and the MLP part of the state adjoint
      {
        float gz[G], bsum = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          gz[g] = gh[g] * act_grad_from_out(c.hidden_act, hreg[g][0]);
          bsum += gz[g];
        }
        acc[u.a_b0 + lane] += bsum;
        row_axpy_g<G>(acc + u.a_w0 + lane * u.s0, x, XR, c.IN4 / 4, gz);
        if (!c.detach_input) {
#pragma unroll
          for (int g = 0; g < G; ++g) gzs[g * H + lane] = gz[g];
          __syncwarp();
          float gx[G];
          row_dot_g<G>(Ws + u.w0t + lane * HS, gzs, H, H / 4, 0.f, gx);
          if (lane < c.IN) {
#pragma unroll
            for (int g = 0; g < G; ++g) gr[g * XR + lane] += gx[g];
          }
        }
      }
      __syncwarp();
    }
  }
  // ---- this warp's partial gradient slab in state_dict layout
  float* out = partials + static_cast<int64_t>(gwarp) * p_stride;
  __syncwarp();
  {
    const int n_out = c.w[1], n_in = c.IN;
    for (int i = lane; i < n_out * n_in; i += 32) out[c.gw[0] + i] = acc[u.a_w0 + (i / n_in) * u.s0 + (i % n_in)];
    for (int i = lane; i < n_out; i += 32) out[c.gb[0] + i] = acc[u.a_b0 + i];
  }
  for (int l = 0; l < c.NHH; ++l) {
    const int n_out = c.w[l + 2], n_in = c.w[l + 1];
    for (int i = lane; i < n_out * n_in; i += 32) out[c.gw[l + 1] + i] = acc[u.a_wh[l] + (i / n_in) * HS + (i % n_in)];
    for (int i = lane; i < n_out; i += 32) out[c.gb[l + 1] + i] = acc[u.a_bh[l] + i];
  }
  {
    const int n_out = c.OUT, n_in = c.w[c.NHH + 1];
    for (int i = lane; i < n_out * n_in; i += 32)
      out[c.gw[c.NHH + 1] + i] = acc[u.a_wo + (i / n_in) * HS + (i % n_in)];
    for (int i = lane; i < n_out; i += 32) out[c.gb[c.NHH + 1] + i] = acc[u.a_bo + i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int unit_sm_count() {
#ifdef HDPO_EMU
  return 1;
#else
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
#endif
}

static int g_unit_max = -1;
void set_unit_max_batch(int max_b) { g_unit_max = max_b < 0 ? -1 : max_b; }
static int g_unit_g = -1;  // scenarios per warp: 1 / 2 / 4, 0 = by batch size (HDPO_SMALL_UNIT_G)
void set_unit_group(int g) { g_unit_g = (g == 1 || g == 2 || g == 4) ? g : 0; }
// Batches up to this many scenarios take the lane = unit kernels (HDPO_SMALL_UNIT_MAX; 0 = never). With one scenario
// per warp the form executes ~5x the warp instructions per scenario of the lane = scenario kernels (517 vs 94 per
// scenario-period forward) but has 32x the independent warps, so it wins while the other form cannot fill the machine.
// Measured on B200, fwd + adjoint of 50 periods, ms per step unit (G = 1) / scenario form:  one-store 1024: 0.49 / 1.06,
// 2048: 0.62 / 1.08, 4096: 0.84 / 1.11, 6144: 1.12 / 1.12, 8192: 1.39 / 1.12;  serial 1024: 0.70 / 0.98,
// 2048: 0.83 / 1.00, 4096: 1.11 / 1.04.
static int unit_max_batch(const Cfg& c) {
  if (g_unit_max < 0) {
    const char* e = getenv("HDPO_SMALL_UNIT_MAX");
    g_unit_max = e ? atoi(e) : -2;
  }
  if (g_unit_max == -2) return c.arch == HDPO_ARCH_VANILLA_ONE_STORE ? kUnitMaxOneStore : kUnitMaxSerial;
  return g_unit_max;
}
static int unit_group(const Cfg& c) {
  if (g_unit_g < 0) {
    const char* e = getenv("HDPO_SMALL_UNIT_G");
    set_unit_group(e ? atoi(e) : 0);
  }
  if (g_unit_g > 0) return g_unit_g;
  (void)c;
  return 1;
}
bool use_unit(const Cfg& c) { return c.ckpt == 1 && c.B <= unit_max_batch(c); }

template <int G>
static int forward_unit_g(const Cfg& c, const UnitLayout& u, const float* params, const float* demands,
                          const HdpoStatics* st, const HdpoState* init, float* cost_b, float* report_b,
                          float* reward_tb, float* tape, const HdpoState& fin, void* stream) {
  constexpr int per_warp = G * (XR + HSZ + kMaxOut);
  int wpc = kUnitWarpsFwd;
  const int sms = unit_sm_count();
  const int n_tiles = ceil_div(c.B, G);
  while (wpc > 1 && ceil_div(n_tiles, wpc) < sms) wpc >>= 1;  // few scenarios: smaller CTAs on more SMs
  const size_t smem = static_cast<size_t>(u.total + wpc * per_warp) * sizeof(float);
  const int ctas_needed = ceil_div(n_tiles, wpc);
  const int grid = ctas_needed < sms ? ctas_needed : sms;
#define HDPO_UNIT_FWD(ARCH)                                                                                           \
  do {                                                                                                                \
    auto k = small_unit_fwd_kernel<ARCH, G>;                                                                          \
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,                                 \
                                      static_cast<int>((u.total + kUnitWarpsFwd * per_warp) * sizeof(float))));       \
    HDPO_LAUNCH(k, grid, wpc * 32, smem, stream, c, u, params, demands, *st, *init, cost_b, report_b, reward_tb, tape, \
                fin);                                                                                                 \
  } while (0)
  if (c.arch == HDPO_ARCH_VANILLA_ONE_STORE) HDPO_UNIT_FWD(HDPO_ARCH_VANILLA_ONE_STORE);
  else HDPO_UNIT_FWD(HDPO_ARCH_VANILLA_SERIAL);
#undef HDPO_UNIT_FWD
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

int forward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const HdpoState* init,
                 float* cost_b, float* report_b, float* reward_tb, float* tape, const HdpoState& fin, void* stream) {
  const UnitLayout u = unit_layout(c);
  switch (unit_group(c)) {
    case 4: return forward_unit_g<4>(c, u, params, demands, st, init, cost_b, report_b, reward_tb, tape, fin, stream);
    case 2: return forward_unit_g<2>(c, u, params, demands, st, init, cost_b, report_b, reward_tb, tape, fin, stream);
    default: return forward_unit_g<1>(c, u, params, demands, st, init, cost_b, report_b, reward_tb, tape, fin, stream);
  }
}

template <int G>
static int backward_unit_g(const Cfg& c, const UnitLayout& u, const float* params, const float* demands,
                           const HdpoStatics* st, const float* tape, float g_total, float g_report, float* partials,
                           int p_stride, int* n_rows, void* stream) {
  const int per_warp = u.a_total + G * (2 * XR + HSZ + H + kMaxOut);
  int wpc_max = kUnitWarpsBwd;
  while (wpc_max > 1 && static_cast<size_t>(u.total + wpc_max * per_warp) * sizeof(float) > 227 * 1024) --wpc_max;
  int wpc = wpc_max;
  const int sms = unit_sm_count();
  const int n_tiles = ceil_div(c.B, G);
  while (wpc > 1 && ceil_div(n_tiles, wpc) < sms) wpc >>= 1;
  const size_t smem = static_cast<size_t>(u.total + wpc * per_warp) * sizeof(float);
  const int ctas_needed = ceil_div(n_tiles, wpc);
  int grid = ctas_needed < sms ? ctas_needed : sms;
  if (grid * wpc > kMaxPartialRows) grid = kMaxPartialRows / wpc;
#define HDPO_UNIT_BWD(ARCH)                                                                                           \
  do {                                                                                                                \
    auto k = small_unit_bwd_kernel<ARCH, G>;                                                                          \
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,                                 \
                                      static_cast<int>((u.total + wpc_max * per_warp) * sizeof(float))));             \
    HDPO_LAUNCH(k, grid, wpc * 32, smem, stream, c, u, params, demands, *st, tape, g_total, g_report, partials,       \
                p_stride);                                                                                            \
  } while (0)
  if (c.arch == HDPO_ARCH_VANILLA_ONE_STORE) HDPO_UNIT_BWD(HDPO_ARCH_VANILLA_ONE_STORE);
  else HDPO_UNIT_BWD(HDPO_ARCH_VANILLA_SERIAL);
#undef HDPO_UNIT_BWD
  HDPO_LAUNCH_OK();
  *n_rows = grid * wpc;
  return HDPO_OK;
}

// *n_rows = the number of partial slabs written (rows of `partials`)
int backward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
                  float g_total, float g_report, float* partials, int p_stride, int* n_rows, void* stream) {
  const UnitLayout u = unit_layout(c);
  switch (unit_group(c)) {
    case 4: return backward_unit_g<4>(c, u, params, demands, st, tape, g_total, g_report, partials, p_stride, n_rows, stream);
    case 2: return backward_unit_g<2>(c, u, params, demands, st, tape, g_total, g_report, partials, p_stride, n_rows, stream);
    default: return backward_unit_g<1>(c, u, params, demands, st, tape, g_total, g_report, partials, p_stride, n_rows, stream);
  }
}

}  // namespace small
}  // namespace hdpo

// Largest batch (scenarios) the lane = unit small-net kernels take: > 0 sets it, 0 = never, < 0 = default.
extern "C" int hdpo_debug_set_small_unit(int32_t max_batch) {
  hdpo::small::set_unit_max_batch(max_batch);
  return HDPO_OK;
}
// Scenarios per warp of the lane = unit kernels: 1 / 2 / 4, anything else = chosen by the batch size.
extern "C" int hdpo_debug_set_small_unit_group(int32_t g) {
  hdpo::small::set_unit_group(g);
  return HDPO_OK;
}
