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

// sum_k row[k] * vec[k] over K4 float4 chunks: this lane's own row against a vector every lane reads (broadcast)
__device__ __forceinline__ float row_dot(const float* __restrict__ row, const float* __restrict__ vec, int K4, float z) {
  float s0 = z, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
  for (int k4 = 0; k4 < K4; ++k4) {
    const float4 w = reinterpret_cast<const float4*>(row)[k4];
    const float4 v = reinterpret_cast<const float4*>(vec)[k4];
    s0 = fmaf(w.x, v.x, s0);
    s1 = fmaf(w.y, v.y, s1);
    s2 = fmaf(w.z, v.z, s2);
    s3 = fmaf(w.w, v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
// row[k] += g * vec[k] over K4 float4 chunks (this lane's accumulator row)
__device__ __forceinline__ void row_axpy(float* __restrict__ row, const float* __restrict__ vec, int K4, float g) {
#pragma unroll 8
  for (int k4 = 0; k4 < K4; ++k4) {
    float4 a = reinterpret_cast<float4*>(row)[k4];
    const float4 v = reinterpret_cast<const float4*>(vec)[k4];
    a.x = fmaf(g, v.x, a.x);
    a.y = fmaf(g, v.y, a.y);
    a.z = fmaf(g, v.z, a.z);
    a.w = fmaf(g, v.w, a.w);
    reinterpret_cast<float4*>(row)[k4] = a;
  }
}

// MLP forward of the warp's scenario: x = state row (IN4 floats), hs = (NHH + 1) activation vectors of H floats.
// hreg[l] = this lane's unit of hidden layer l; y = the outputs (every lane gets all of them).
__device__ __forceinline__ void unit_mlp_fwd(const Cfg& c, const UnitLayout& u, const float* __restrict__ Ws,
                                             const float* __restrict__ x, float* __restrict__ hs, int lane,
                                             float (&hreg)[kMaxHH + 1], float (&y)[kMaxOut]) {
  float z = row_dot(Ws + u.w0n + lane * u.s0, x, c.IN4 / 4, Ws[u.b0 + lane]);
  float h = act_fwd(c.hidden_act, z);
  hreg[0] = h;
  hs[lane] = h;
  __syncwarp();
#pragma unroll
  for (int l = 0; l < kMaxHH; ++l) {
    if (l < c.NHH) {
      z = row_dot(Ws + u.whn[l] + lane * HS, hs + l * H, H / 4, Ws[u.bh[l] + lane]);
      h = act_fwd(c.hidden_act, z);
      hreg[l + 1] = h;
      hs[(l + 1) * H + lane] = h;
      __syncwarp();
    }
  }
  const int o = lane < c.OUT ? lane : 0;
  const float yo = row_dot(Ws + u.wo + o * HS, hs + c.NHH * H, H / 4, Ws[u.bo + o]);
#pragma unroll
  for (int i = 0; i < kMaxOut; ++i) {
    const float v = __shfl_sync(0xffffffffu, yo, i);
    y[i] = i < c.OUT ? v : 0.f;
  }
}

__device__ __forceinline__ void unit_load_state(const Cfg& c, float* __restrict__ x, const float* __restrict__ src, int lane) {
  if (lane < c.IN4 / 4) reinterpret_cast<float4*>(x)[lane] = reinterpret_cast<const float4*>(src)[lane];
}

template <int ARCH>
__global__ void __launch_bounds__(kUnitWarpsFwd * 32, 1)
small_unit_fwd_kernel(Cfg c, UnitLayout u, const float* __restrict__ params, const float* __restrict__ demands,
                      HdpoStatics st, HdpoState init, float* __restrict__ cost_b, float* __restrict__ report_b,
                      float* __restrict__ reward_tb, float* __restrict__ tape, HdpoState fin) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights_unit(c, u, params, Ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int per_warp = kMaxIn + 4 + (kMaxHH + 1) * H;
  float* x = smem + u.total + warp * per_warp;
  float* hs = x + kMaxIn + 4;
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  for (int b = gwarp; b < c.B; b += nwarps) {
    Statics s;
    load_statics<ARCH>(c, st, b, s);
    __syncwarp();
    // initial state row: [store L | warehouse Lw | echelons E*Le], zero padded to IN4
    if (lane < c.IN4) {
      float v = 0.f;
      if (lane < c.L) v = init.store[static_cast<int64_t>(b) * c.L + lane];
      else if (lane < c.L + c.Lw) v = init.warehouse[static_cast<int64_t>(b) * c.Lw + (lane - c.L)];
      else if (lane < c.IN) v = init.echelon[static_cast<int64_t>(b) * c.E * c.Le + (lane - c.L - c.Lw)];
      x[lane] = v;
    }
    __syncwarp();
    float cost = 0.f, rep = 0.f;
    float dnext = demand_at(c, demands, b, 0);
    for (int t = 0; t < c.T; ++t) {
      const float d = dnext;
      if (t + 1 < c.T) dnext = demand_at(c, demands, b, t + 1);
      if (tape && t % c.ckpt == 0 && lane < c.IN4 / 4)
        reinterpret_cast<float4*>(tape + (static_cast<int64_t>(t / c.ckpt) * c.B + b) * c.tape_stride)[lane] =
            reinterpret_cast<const float4*>(x)[lane];
      float hreg[kMaxHH + 1], y[kMaxOut];
      unit_mlp_fwd(c, u, Ws, x, hs, lane, hreg, y);
      if (lane == 0) {
        Head hd;
        head_fwd<ARCH>(c, x, y, hd);
        const float r = env_fwd<ARCH>(c, x, d, hd, s);
        cost += r;
        if (t >= c.ignore) rep += r;
        if (reward_tb) reward_tb[static_cast<int64_t>(t) * c.B + b] = r;
      }
      __syncwarp();
    }
    if (lane == 0) {
      cost_b[b] = cost;
      if (report_b) report_b[b] = rep;
    }
    if (fin.store && lane < c.L) fin.store[static_cast<int64_t>(b) * c.L + lane] = x[lane];
    if (fin.warehouse && lane >= c.L && lane < c.L + c.Lw)
      fin.warehouse[static_cast<int64_t>(b) * c.Lw + (lane - c.L)] = x[lane];
    if (fin.echelon && lane >= c.L + c.Lw && lane < c.IN)
      fin.echelon[static_cast<int64_t>(b) * c.E * c.Le + (lane - c.L - c.Lw)] = x[lane];
  }
}

template <int ARCH>
__global__ void __launch_bounds__(kUnitWarpsBwd * 32, 1)
small_unit_bwd_kernel(Cfg c, UnitLayout u, const float* __restrict__ params, const float* __restrict__ demands,
                      HdpoStatics st, const float* __restrict__ tape, float g_total, float g_report,
                      float* __restrict__ partials, int p_stride) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights_unit(c, u, params, Ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int per_warp = u.a_total + 2 * (kMaxIn + 4) + (kMaxHH + 1) * H + H + kMaxOut;
  float* acc = smem + u.total + warp * per_warp;
  float* x = acc + u.a_total;
  float* g = x + kMaxIn + 4;
  float* hs = g + kMaxIn + 4;
  float* gzs = hs + (kMaxHH + 1) * H;
  float* gys = gzs + H;
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  for (int i = lane; i < u.a_total; i += 32) acc[i] = 0.f;
  __syncwarp();
  for (int b = gwarp; b < c.B; b += nwarps) {
    Statics s;
    load_statics<ARCH>(c, st, b, s);
    __syncwarp();
    g[lane] = 0.f;  // adjoint wrt the state after the last period
    if (lane < 4) g[32 + lane] = 0.f;
    for (int t = c.T - 1; t >= 0; --t) {
      unit_load_state(c, x, tape + (static_cast<int64_t>(t) * c.B + b) * c.tape_stride, lane);
      const float d = demand_at(c, demands, b, t);
      __syncwarp();
      float hreg[kMaxHH + 1], y[kMaxOut];
      unit_mlp_fwd(c, u, Ws, x, hs, lane, hreg, y);
      if (lane == 0) {
        Head hd;
        head_fwd<ARCH>(c, x, y, hd);
        const float rb = g_total + (t >= c.ignore ? g_report : 0.f);
        float gy[kMaxOut];
        head_env_bwd<ARCH>(c, x, g, d, hd, s, rb, gy);
#pragma unroll
        for (int i = 0; i < kMaxOut; ++i) gys[i] = gy[i];
      }
      __syncwarp();
      // ---- output layer: dWo[o][k] += gy[o] h[k] (lane = k), dbo, gh[k] = sum_o Wo[o][k] gy[o]
      float gh = 0.f;
      {
        float hl = hreg[0];
#pragma unroll
        for (int l = 1; l <= kMaxHH; ++l)
          if (l == c.NHH) hl = hreg[l];
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o) {
          if (o < c.OUT) {
            const float gyo = gys[o];
            acc[u.a_wo + o * HS + lane] = fmaf(gyo, hl, acc[u.a_wo + o * HS + lane]);
            gh = fmaf(Ws[u.wo + o * HS + lane], gyo, gh);
            if (lane == o) acc[u.a_bo + o] += gyo;
          }
        }
      }
      // ---- hidden layers, last to first: gz = gh * act'(h); dW[n][:] += gz[n] h_prev[:]; gh_prev = W^T gz
#pragma unroll
      for (int l = kMaxHH - 1; l >= 0; --l) {
        if (l < c.NHH) {
          const float gz = gh * act_grad_from_out(c.hidden_act, hreg[l + 1]);
          gzs[lane] = gz;
          acc[u.a_bh[l] + lane] += gz;
          row_axpy(acc + u.a_wh[l] + lane * HS, hs + l * H, H / 4, gz);
          __syncwarp();
          gh = row_dot(Ws + u.wht[l] + lane * HS, gzs, H / 4, 0.f);
          __syncwarp();
        }
      }
      // ---- first layer: weight gradient against the state row, and the MLP part of the state adjoint
      {
        const float gz = gh * act_grad_from_out(c.hidden_act, hreg[0]);
        acc[u.a_b0 + lane] += gz;
        row_axpy(acc + u.a_w0 + lane * u.s0, x, c.IN4 / 4, gz);
        if (!c.detach_input) {
          gzs[lane] = gz;
          __syncwarp();
          const float gx = row_dot(Ws + u.w0t + lane * HS, gzs, H / 4, 0.f);
          if (lane < c.IN) g[lane] += gx;
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
// Batches up to this many scenarios take the lane = unit kernels (HDPO_SMALL_UNIT_MAX; 0 = never). The form executes
// ~5x the warp instructions per scenario of the lane = scenario kernels (517 vs 94 per scenario-period forward) but has
// 32x the independent warps, so it wins while the other form cannot fill the machine. Measured on B200, fwd + adjoint
// of 50 periods, ms per step unit / scenario form:  one-store 1024: 0.49 / 1.06, 2048: 0.62 / 1.08, 4096: 0.84 / 1.11,
// 6144: 1.12 / 1.12, 8192: 1.39 / 1.12;  serial 1024: 0.70 / 0.98, 2048: 0.83 / 1.00, 4096: 1.11 / 1.04.
static int unit_max_batch(const Cfg& c) {
  if (g_unit_max < 0) {
    const char* e = getenv("HDPO_SMALL_UNIT_MAX");
    g_unit_max = e ? atoi(e) : -2;
  }
  if (g_unit_max == -2) return c.arch == HDPO_ARCH_VANILLA_ONE_STORE ? 4096 : 2048;
  return g_unit_max;
}
bool use_unit(const Cfg& c) { return c.ckpt == 1 && c.B <= unit_max_batch(c); }

int forward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const HdpoState* init,
                 float* cost_b, float* report_b, float* reward_tb, float* tape, const HdpoState& fin, void* stream) {
  const UnitLayout u = unit_layout(c);
  const int per_warp = kMaxIn + 4 + (kMaxHH + 1) * H;
  int wpc = kUnitWarpsFwd;
  const int sms = unit_sm_count();
  while (wpc > 1 && ceil_div(c.B, wpc) < sms) wpc >>= 1;  // few scenarios: smaller CTAs on more SMs
  const size_t smem = static_cast<size_t>(u.total + wpc * per_warp) * sizeof(float);
  const int ctas_needed = ceil_div(c.B, wpc);
  const int grid = ctas_needed < sms ? ctas_needed : sms;
#define HDPO_UNIT_FWD(ARCH)                                                                                           \
  do {                                                                                                                \
    auto k = small_unit_fwd_kernel<ARCH>;                                                                             \
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

// *n_rows = the number of partial slabs written (rows of `partials`)
int backward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
                  float g_total, float g_report, float* partials, int p_stride, int* n_rows, void* stream) {
  const UnitLayout u = unit_layout(c);
  const int per_warp = u.a_total + 2 * (kMaxIn + 4) + (kMaxHH + 1) * H + H + kMaxOut;
  int wpc = kUnitWarpsBwd;
  const int sms = unit_sm_count();
  while (wpc > 1 && ceil_div(c.B, wpc) < sms) wpc >>= 1;
  const size_t smem = static_cast<size_t>(u.total + wpc * per_warp) * sizeof(float);
  const int ctas_needed = ceil_div(c.B, wpc);
  int grid = ctas_needed < sms ? ctas_needed : sms;
  if (grid * wpc > kMaxPartialRows) grid = kMaxPartialRows / wpc;
#define HDPO_UNIT_BWD(ARCH)                                                                                           \
  do {                                                                                                                \
    auto k = small_unit_bwd_kernel<ARCH>;                                                                             \
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,                                 \
                                      static_cast<int>((u.total + kUnitWarpsBwd * per_warp) * sizeof(float))));       \
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

}  // namespace small
}  // namespace hdpo

// Largest batch (scenarios) the lane = unit small-net kernels take: > 0 sets it, 0 = never, < 0 = default.
extern "C" int hdpo_debug_set_small_unit(int32_t max_batch) {
  hdpo::small::set_unit_max_batch(max_batch);
  return HDPO_OK;
}
