// rollout_sym.cu - per-scenario part of the fused SymmetryAware rollout (see rollout_sym.cuh; policy structure from
// SURVEY.md 2.3, simulator period from environment.py:110-270, proportional allocation neural_networks.py:111-138).
//
// Design: one WARP per scenario, persistent over the scenarios of the launch. Inside a scenario
//   * the store net (same weights for every store) runs with lane = store, 32 (x NS) stores at a time, on the
//     small-net machinery of rollout_small_kernels.cuh: weights staged once per CTA as Wt[k][n] so one broadcast
//     LDS.128 feeds 4 (x NS) FFMAs of 32 independent accumulators per lane; the first layer starts from the
//     projection row (context half of the layer + bias, computed by the trunk GEMM once per scenario instead of once
//     per store: the factored form of SURVEY.md 8d);
//   * the warehouse net (one row per scenario) runs with lane = hidden unit;
//   * proportional allocation, the store / warehouse period and the cost are warp reductions over the store lanes.
// The adjoint recomputes the local nets (only the store outputs are taped: S floats per scenario-period), keeps the
// store-net weight gradients in registers as per-lane 4x8 tiles (wgrad_tile) and the warehouse-net ones in shared
// memory, and adds them to a per-warp slab in global memory at the end of each launch; the slabs are reduced in a
// fixed order once per batch, so the gradient is deterministic.
#include "rollout_sym.cuh"

#include <cstdlib>

#include "mma32.cuh"
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace sym {

using small::H;
using small::HS;
constexpr int WS = 33;  // warehouse-net weight rows: [k][n] with stride 33, conflict-free for lane = n AND lane = k
constexpr int kMaxWarps = 8;
constexpr size_t kSmemMax = 220 * 1024;

static inline int pad_to(int x, int q) { return (x + q - 1) / q * q; }

static int mlp_params(const HdpoMlp& m) {
  int n = 0;
  for (int i = 0; i < m.n_layers; ++i) n += m.widths[i + 1] * m.widths[i] + m.widths[i + 1];
  return n;
}

static bool local_net_ok(const HdpoMlp& m, int n_in) {
  if (m.n_layers < 2 || m.n_layers > kMaxHH + 2) return false;
  if (m.widths[0] != n_in || m.widths[m.n_layers] != 1) return false;
  for (int i = 1; i < m.n_layers; ++i)
    if (m.widths[i] < 1 || m.widths[i] > H) return false;
  return true;
}

bool supported(const HdpoRolloutDesc* d) {
  if (d->arch != HDPO_ARCH_SYMMETRY_AWARE) return false;
  const HdpoProblem& pb = d->pb;
  if (pb.W != 1 || pb.E != 0 || pb.S < 1 || pb.S > 256) return false;
  if (d->transshipment) return false;
  const HdpoMlp& m = d->master;
  if (m.n_layers < 1 || m.n_layers + 1 > HDPO_MAX_LAYERS) return false;
  if (m.widths[0] != pb.S * pb.L + pb.Lw) return false;
  const int C = m.widths[m.n_layers];
  if (pb.L + 4 > 32 || pb.Lw > 32) return false;
  if (!local_net_ok(d->store_net, pb.L + 4 + C) || !local_net_ok(d->warehouse_net, pb.Lw + C)) return false;
  return true;
}

int build_cfg(const HdpoRolloutDesc* d, int B, int Bp, int ldx, int ldy, Cfg* c) {
  const HdpoProblem& pb = d->pb;
  const HdpoMlp &m = d->master, &sn = d->store_net, &wn = d->warehouse_net;
  HDPO_REQUIRE(ldy == 2 * H, "the projection row must be %d floats wide", 2 * H);
  c->B = B;
  c->Bp = Bp;
  c->S = pb.S;
  c->SP = pad_to(pb.S, 32);
  c->L = pb.L;
  c->Lw = pb.Lw;
  c->C = m.widths[m.n_layers];
  c->ldx = ldx;
  c->ldy = ldy;
  c->ldo = c->SP;
  c->nS = pb.S * pb.L;
  c->s_in = pb.L + 4;
  c->s_kq = c->s_in <= 8 ? 2 : (c->s_in <= 16 ? 4 : 8);
  c->s_in4 = 4 * c->s_kq;
  c->s_xs = c->s_in4 + 4;
  c->s_nhh = sn.n_layers - 2;
  c->s_hact = sn.hidden_act;
  c->s_oact = sn.out_act;
  for (int i = 0; i <= c->s_nhh; ++i) c->s_w[i] = sn.widths[i + 1];
  c->s_ld0 = sn.widths[0];
  c->w_nhh = wn.n_layers - 2;
  c->w_hact = wn.hidden_act;
  c->w_oact = wn.out_act;
  for (int i = 0; i <= c->w_nhh; ++i) c->w_w[i] = wn.widths[i + 1];
  c->w_ld0 = wn.widths[0];
  // flat parameter offsets (state_dict order: context, store, warehouse; weight then bias per layer)
  int off = mlp_params(m);
  c->g_s_w0 = off;
  off += c->s_w[0] * c->s_ld0;
  c->g_s_b0 = off;
  off += c->s_w[0];
  for (int l = 0; l < c->s_nhh; ++l) {
    c->g_s_wh[l] = off;
    off += c->s_w[l + 1] * c->s_w[l];
    c->g_s_bh[l] = off;
    off += c->s_w[l + 1];
  }
  c->g_s_wo = off;
  off += c->s_w[c->s_nhh];
  c->g_s_bo = off;
  off += 1;
  c->g_w_w0 = off;
  off += c->w_w[0] * c->w_ld0;
  c->g_w_b0 = off;
  off += c->w_w[0];
  for (int l = 0; l < c->w_nhh; ++l) {
    c->g_w_wh[l] = off;
    off += c->w_w[l + 1] * c->w_w[l];
    c->g_w_bh[l] = off;
    off += c->w_w[l + 1];
  }
  c->g_w_wo = off;
  off += c->w_w[c->w_nhh];
  c->g_w_bo = off;
  off += 1;
  c->P = off;
  // shared-memory weight block
  int s = 0;
  auto take = [&](int n) {
    const int at = s;
    s += pad_to(n, 4);
    return at;
  };
  c->m_s_wt0 = take(c->s_in4 * H);
  for (int l = 0; l < c->s_nhh; ++l) {
    c->m_s_wth[l] = take(H * H);
    c->m_s_bh[l] = take(H);
  }
  c->m_s_wo = take(H);
  c->m_s_bo = take(4);
  c->m_w_wt0 = take(c->Lw * WS);
  for (int l = 0; l < c->w_nhh; ++l) {
    c->m_w_wth[l] = take(H * WS);
    c->m_w_bh[l] = take(H);
  }
  c->m_w_wo = take(H);
  c->m_w_bo = take(4);
  c->m_total_simt = s;  // weight block of a kernel that does not use the mma forms
#ifdef HDPO_EMU
  c->tc = 0;
#else
  c->tc = d->precision != HDPO_PREC_FP32;
#endif
  // Where the mma.sync forms pay (measured, 8192 scenarios x 50 stores): adjoint head 18.3 -> 16.9 ms per step; the
  // forward head is FASTER in the FFMA2 form (6.8 vs 7.1 - 8.0 ms: at 8 warps per SM the HMMA latency is not hidden and
  // two stores per lane amortise the weight loads), so it stays SIMT. HDPO_SYM_TC_FWD / HDPO_SYM_TC_RECOMPUTE override.
  {
    const char* e = getenv("HDPO_SYM_TC_FWD");
    c->tc_fwd = c->tc && e && atoi(e) != 0;
    e = getenv("HDPO_SYM_TC_RECOMPUTE");
    c->tc_recompute = c->tc && (!e || atoi(e) != 0);
  }
  if (c->tc) {
    c->m_s_w0n_hi = take(H * c->s_xs);
    c->m_s_w0n_lo = take(H * c->s_xs);
    for (int l = 0; l < c->s_nhh; ++l) {
      c->m_s_whn_hi[l] = take(H * HS);
      c->m_s_whn_lo[l] = take(H * HS);
      c->m_s_whk_hi[l] = take(H * HS);
      c->m_s_whk_lo[l] = take(H * HS);
    }
  }
  c->m_total = s;
  // gradient slab + parameter blocks
  int q = 0, nb = 0;
  auto block = [&](int* at, int q_ld, int rows, int cols, int dst, int dst_ld, int floats) {
    *at = q;
    c->blk[nb++] = Block{q, q_ld, rows, cols, dst, dst_ld};
    q += pad_to(floats, 4);
  };
  block(&c->q_s_w0, c->s_in4, c->s_w[0], c->s_in, c->g_s_w0, c->s_ld0, H * c->s_in4);
  for (int l = 0; l < c->s_nhh; ++l) {
    block(&c->q_s_wh[l], H, c->s_w[l + 1], c->s_w[l], c->g_s_wh[l], c->s_w[l], H * H);
    block(&c->q_s_bh[l], H, 1, c->s_w[l + 1], c->g_s_bh[l], c->s_w[l + 1], H);
  }
  block(&c->q_s_wo, H, 1, c->s_w[c->s_nhh], c->g_s_wo, c->s_w[c->s_nhh], H);
  block(&c->q_s_bo, 4, 1, 1, c->g_s_bo, 1, 4);
  const int lw4 = pad_to(c->Lw, 4);
  block(&c->q_w_w0, lw4, c->w_w[0], c->Lw, c->g_w_w0, c->w_ld0, H * lw4);
  for (int l = 0; l < c->w_nhh; ++l) {
    block(&c->q_w_wh[l], H, c->w_w[l + 1], c->w_w[l], c->g_w_wh[l], c->w_w[l], H * H);
    block(&c->q_w_bh[l], H, 1, c->w_w[l + 1], c->g_w_bh[l], c->w_w[l + 1], H);
  }
  block(&c->q_w_wo, H, 1, c->w_w[c->w_nhh], c->g_w_wo, c->w_w[c->w_nhh], H);
  block(&c->q_w_bo, 4, 1, 1, c->g_w_bo, 1, 4);
  c->q_total = q;
  c->n_blocks = nb;
  c->lost = pb.lost_demand;
  c->profit = pb.maximize_profit;
  c->has_edge = pb.has_edge_cost;
  c->discrete = d->discrete_allocation;
  c->t_stride = d->t_stride;
  c->demand_layout = d->demand_layout;
  c->demand_bstride = pb.B;
  c->wub = d->warehouse_upper_bound;
  c->eps = d->prop_eps;
  return HDPO_OK;
}

// warehouse-net gradient accumulators of one warp in shared memory (floats)
__host__ __device__ inline int wacc_floats(const Cfg& c) {
  const int lws = c.Lw | 1;
  return H * lws + c.w_nhh * (H * WS + H) + H + 4;
}
// stores per lane of one forward store-net pass (tuning knob: HDPO_SYM_FWD_NS)
static int fwd_ns(const Cfg& c) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("HDPO_SYM_FWD_NS");
    env = e ? atoi(e) : 0;
  }
  if (env == 1 || env == 2) return env;
  return c.S > 32 ? 2 : 1;
}
int fwd_smem_floats_per_warp(const Cfg& c) {
  const int ns = fwd_ns(c);
  const int n = 2 * c.ldx + c.ldy + 7 * c.SP + ns * 32 * (c.s_xs + HS) + (kMaxHH + 1) * H + 32;
  return pad_to(n, 4);
}
int bwd_smem_floats_per_warp(const Cfg& c) {
  const int n = 2 * c.ldx + 2 * c.ldy + 8 * c.SP + 32 * c.s_xs + (c.s_nhh + 1) * 32 * HS + 32 + (kMaxHH + 1) * H + 2 * H +
                pad_to(wacc_floats(c), 4) + 32;
  return pad_to(n, 4);
}

// floats of the shared weight block a head kernel stages: the mma copies only where that kernel uses them
__host__ __device__ inline int weight_floats(const Cfg& c, bool bwd) {
  const bool mma = bwd ? c.tc != 0 : c.tc_fwd != 0;
  return ((mma ? c.m_total : c.m_total_simt) + 3) & ~3;
}

static int warps_per_cta(const Cfg& c, int per_warp_floats, int rows, bool bwd) {
  int w = kMaxWarps;
  {
    static int env = -1;
    if (env < 0) {
      const char* e = getenv("HDPO_SYM_WPC");
      env = e ? atoi(e) : 0;
    }
    if (env >= 1 && env <= kMaxWarps) w = env;
  }
  while (w > 1 && (static_cast<size_t>(weight_floats(c, bwd)) + static_cast<size_t>(w) * per_warp_floats) * sizeof(float) > kSmemMax) --w;
  while (w > 1 && rows < w * 8) w >>= 1;  // few scenarios: more, smaller CTAs
  return w;
}

static int sm_count() {
#ifdef HDPO_EMU
  return 2;
#else
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
#endif
}

static size_t smem_bytes(const Cfg& c, int wpc, int per_warp_floats, bool bwd) {
  return (static_cast<size_t>(weight_floats(c, bwd)) + static_cast<size_t>(wpc) * per_warp_floats) * sizeof(float);
}
// persistent grid: at most the CTAs that are resident at once (shared memory decides how many fit on an SM)
static int max_resident_ctas(size_t smem) {
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  return sm_count() * per_sm;
}

// launch shape of the adjoint head: fixed by (B, shapes) so that warp w owns the same scenarios in every period
static void bwd_shape(const Cfg& c, int* grid, int* wpc) {
  const int per_warp = bwd_smem_floats_per_warp(c);
  *wpc = warps_per_cta(c, per_warp, c.Bp, true);
  // one scenario per warp until the machine is full, more beyond that (the per-launch slab update of a warp, ~4 k
  // floats through L2, is ~10 % of one scenario's work; at 1024 scenarios four-per-warp left 116 SMs idle: adjoint
  // sweep 8.75 ms per step)
  *grid = ceil_div(c.Bp, *wpc);
  const int cap = max_resident_ctas(smem_bytes(c, *wpc, per_warp, true));
  if (*grid > cap) *grid = cap;
}
int bwd_warps(const Cfg& c) {
  int grid, wpc;
  bwd_shape(c, &grid, &wpc);
  return grid * wpc;
}

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float tf32_round(float x) {
#ifdef HDPO_EMU
  return x;
#else
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#endif
}

__device__ __forceinline__ float demand_of(const Cfg& c, const PeriodArgs& a, int b, int s) {
  if (c.demand_layout == HDPO_DEMAND_TSB)
    return __ldg(a.demands + (static_cast<size_t>(a.tt) * c.S + s) * c.demand_bstride + b);
  return __ldg(a.demands + (static_cast<size_t>(b) * c.S + s) * c.t_stride + a.tt);
}

// flat parameter vector -> shared weight block (zero padded). Store net: Wt[k][n] stride H; warehouse net: stride WS.
static __device__ void stage_weights(const Cfg& c, const float* __restrict__ params, float* __restrict__ Ws, bool bwd) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n_floats = weight_floats(c, bwd);
  const bool mma = n_floats > ((c.m_total_simt + 3) & ~3);
  for (int i = tid; i < n_floats; i += nt) Ws[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < c.s_w[0] * c.s_in; i += nt) {
    const int n = i / c.s_in, k = i % c.s_in;
    Ws[c.m_s_wt0 + k * H + n] = params[c.g_s_w0 + n * c.s_ld0 + k];
  }
  for (int l = 0; l < c.s_nhh; ++l) {
    const int n_out = c.s_w[l + 1], n_in = c.s_w[l];
    for (int i = tid; i < n_out * n_in; i += nt) {
      const int n = i / n_in, k = i % n_in;
      Ws[c.m_s_wth[l] + k * H + n] = params[c.g_s_wh[l] + i];
    }
    for (int i = tid; i < n_out; i += nt) Ws[c.m_s_bh[l] + i] = params[c.g_s_bh[l] + i];
  }
  for (int i = tid; i < c.s_w[c.s_nhh]; i += nt) Ws[c.m_s_wo + i] = params[c.g_s_wo + i];
  if (tid == 0) Ws[c.m_s_bo] = params[c.g_s_bo];
#ifndef HDPO_EMU
  if (mma) {  // (hi, lo) copies for the tensor-core fragments; padding stays zero
    for (int i = tid; i < c.s_w[0] * c.s_in; i += nt) {
      const int n = i / c.s_in, k = i % c.s_in;
      unsigned hi, lo;
      mma32::split(params[c.g_s_w0 + n * c.s_ld0 + k], hi, lo);
      Ws[c.m_s_w0n_hi + n * c.s_xs + k] = __uint_as_float(hi);
      Ws[c.m_s_w0n_lo + n * c.s_xs + k] = __uint_as_float(lo);
    }
    for (int l = 0; l < c.s_nhh; ++l) {
      const int n_out = c.s_w[l + 1], n_in = c.s_w[l];
      for (int i = tid; i < n_out * n_in; i += nt) {
        const int n = i / n_in, k = i % n_in;
        unsigned hi, lo;
        mma32::split(params[c.g_s_wh[l] + i], hi, lo);
        Ws[c.m_s_whn_hi[l] + n * HS + k] = __uint_as_float(hi);
        Ws[c.m_s_whn_lo[l] + n * HS + k] = __uint_as_float(lo);
        Ws[c.m_s_whk_hi[l] + k * HS + n] = __uint_as_float(hi);
        Ws[c.m_s_whk_lo[l] + k * HS + n] = __uint_as_float(lo);
      }
    }
  }
#endif
  for (int i = tid; i < c.w_w[0] * c.Lw; i += nt) {
    const int n = i / c.Lw, k = i % c.Lw;
    Ws[c.m_w_wt0 + k * WS + n] = params[c.g_w_w0 + n * c.w_ld0 + k];
  }
  for (int l = 0; l < c.w_nhh; ++l) {
    const int n_out = c.w_w[l + 1], n_in = c.w_w[l];
    for (int i = tid; i < n_out * n_in; i += nt) {
      const int n = i / n_in, k = i % n_in;
      Ws[c.m_w_wth[l] + k * WS + n] = params[c.g_w_wh[l] + i];
    }
    for (int i = tid; i < n_out; i += nt) Ws[c.m_w_bh[l] + i] = params[c.g_w_bh[l] + i];
  }
  for (int i = tid; i < c.w_w[c.w_nhh]; i += nt) Ws[c.m_w_wo + i] = params[c.g_w_wo + i];
  if (tid == 0) Ws[c.m_w_bo] = params[c.g_w_bo];
  __syncthreads();
}

// local input row of store s: [L pipeline slots | mean | std | underage | lead time | 0 pad]
__device__ __forceinline__ void build_loc_row(const Cfg& c, float* __restrict__ row, const float* __restrict__ xs, int s,
                                              const float* mean, const float* sd, const float* p, const float* lt) {
  for (int k = 0; k < c.L; ++k) row[k] = xs[s * c.L + k];
  row[c.L + 0] = mean[s];
  row[c.L + 1] = sd[s];
  row[c.L + 2] = p[s];
  row[c.L + 3] = lt[s];
  for (int k = c.s_in; k < c.s_in4; ++k) row[k] = 0.f;
}

__device__ __forceinline__ float out_dot(const float* __restrict__ wo, const float* __restrict__ hrow) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < H / 4; ++k4) {
    const float4 wv = reinterpret_cast<const float4*>(wo)[k4];
    const float4 hv = reinterpret_cast<const float4*>(hrow)[k4];
    s0 = fmaf(wv.x, hv.x, s0);
    s1 = fmaf(wv.y, hv.y, s1);
    s2 = fmaf(wv.z, hv.z, s2);
    s3 = fmaf(wv.w, hv.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// apply the hidden activation in place to NS rows of H floats. A short rolled loop on purpose: the heads run once
// per scenario through a long instruction stream, so unrolled activation code (32 elements x NS rows x call sites)
// overflowed the 32 KB instruction cache (ncu: no_instruction was the top stall); the extra LDS/STS pair per float4
// is noise next to the 256 LDS.128 of a layer.
template <int NS>
__device__ __forceinline__ void act_rows_inplace(int act, float* const (&row)[NS]) {
  dispatch_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll 1
    for (int n4 = 0; n4 < H / 4; ++n4) {
      float v[4 * NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const float4 t = reinterpret_cast<float4*>(row[j])[n4];
        v[4 * j + 0] = t.x;
        v[4 * j + 1] = t.y;
        v[4 * j + 2] = t.z;
        v[4 * j + 3] = t.w;
      }
      if (ACT == HDPO_ACT_ELU) {
        elu_inplace(v);  // branch-free: the per-element `x > 0 ? :` form costs a divergent region per element
      } else {
#pragma unroll
        for (int i = 0; i < 4 * NS; ++i) v[i] = act_fwd_t<ACT>(v[i]);
      }
#pragma unroll
      for (int j = 0; j < NS; ++j)
        reinterpret_cast<float4*>(row[j])[n4] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  });
}

// store net for NS rows of this lane. hid[j] + l * layer_stride receives the activations of hidden layer l
// (layer_stride = 0: in place, forward). Returns the pre-activation outputs. ONE call site of the layer code inside a
// rolled loop over the layers (instruction-cache footprint, see act_rows_inplace).
template <int NS>
__device__ __forceinline__ void store_net_fwd(const Cfg& c, const float* __restrict__ Ws, const float* __restrict__ prj,
                                              const float* const (&loc)[NS], float* const (&hid)[NS], int layer_stride,
                                              float (&y)[NS], bool use_mma) {
#pragma unroll 1
  for (int l = 0; l <= c.s_nhh; ++l) {
    const float* Wt = Ws + (l == 0 ? c.m_s_wt0 : c.m_s_wth[l - 1]);
    const float* bias = l == 0 ? prj : Ws + c.m_s_bh[l - 1];
    const int K4 = l == 0 ? c.s_in4 / 4 : H / 4;
    const float* in[NS];
    float* out[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      in[j] = l == 0 ? loc[j] : hid[j] + (l - 1) * layer_stride;
      out[j] = hid[j] + l * layer_stride;
    }
#ifndef HDPO_EMU
    if (use_mma) {
      // tensor-core form: the 32 NS rows of the warp as one M = 32 NS product (rows of block j follow block j - 1)
      const int lane = threadIdx.x & 31;
      const int in_stride = l == 0 ? c.s_xs : HS;
      __syncwarp();  // the input rows were written lane-by-lane
      mma32::layer<2 * NS>(Ws + (l == 0 ? c.m_s_w0n_hi : c.m_s_whn_hi[l - 1]), Ws + (l == 0 ? c.m_s_w0n_lo : c.m_s_whn_lo[l - 1]),
                           in_stride, bias, in[0] - lane * in_stride, in_stride, l == 0 ? c.s_in4 / 8 : H / 8,
                           out[0] - lane * HS, HS, lane);
    } else
#endif
    {
      float2 acc[NS][H / 2];
      small::layer_fwd<NS>(Wt, bias, K4, in, acc);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
#pragma unroll
        for (int n4 = 0; n4 < H / 4; ++n4)
          reinterpret_cast<float4*>(out[j])[n4] =
              make_float4(acc[j][2 * n4].x, acc[j][2 * n4].y, acc[j][2 * n4 + 1].x, acc[j][2 * n4 + 1].y);
      }
    }
    act_rows_inplace<NS>(c.s_hact, out);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j) y[j] = Ws[c.m_s_bo] + out_dot(Ws + c.m_s_wo, hid[j] + c.s_nhh * layer_stride);
}

// warehouse net, lane = unit. hw[l*H + n] receives the activations of hidden layer l; returns the pre-activation
// output (same value in every lane). prj_w = projection row + H (context half of layer 0 + bias).
__device__ __forceinline__ float wh_net_fwd(const Cfg& c, const float* __restrict__ Ws, const float* __restrict__ prj_w,
                                            const float* __restrict__ xw, float* __restrict__ hw, int lane) {
  float z = prj_w[lane];
  for (int k = 0; k < c.Lw; ++k) z = fmaf(Ws[c.m_w_wt0 + k * WS + lane], xw[k], z);
  hw[lane] = act_fwd(c.w_hact, z);
  __syncwarp();
  for (int l = 0; l < c.w_nhh; ++l) {
    z = Ws[c.m_w_bh[l] + lane];
    const float* Wt = Ws + c.m_w_wth[l];
#pragma unroll 8
    for (int k = 0; k < H; ++k) z = fmaf(Wt[k * WS + lane], hw[l * H + k], z);
    hw[(l + 1) * H + lane] = act_fwd(c.w_hact, z);
    __syncwarp();
  }
  return Ws[c.m_w_bo] + warp_sum(Ws[c.m_w_wo + lane] * hw[c.w_nhh * H + lane]);
}

__device__ __forceinline__ void store_row_split(float* __restrict__ dst, float* __restrict__ dst_hi,
                                                float* __restrict__ dst_lo, const float* __restrict__ src, int n, int lane) {
  for (int k = lane * 4; k < n; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(src + k);
    if (dst) *reinterpret_cast<float4*>(dst + k) = v;
    if (dst_hi) {
      float4 hi, lo;
      hi.x = tf32_round(v.x);
      hi.y = tf32_round(v.y);
      hi.z = tf32_round(v.z);
      hi.w = tf32_round(v.w);
      lo.x = tf32_round(v.x - hi.x);
      lo.y = tf32_round(v.y - hi.y);
      lo.z = tf32_round(v.z - hi.z);
      lo.w = tf32_round(v.w - hi.w);
      *reinterpret_cast<float4*>(dst_hi + k) = hi;
      *reinterpret_cast<float4*>(dst_lo + k) = lo;
    }
  }
}

// per-warp shared-memory rows common to both heads
struct Rows {
  float *xs, *xo, *prj, *so, *lt, *h, *p, *d, *mean, *sd;
};
__device__ __forceinline__ float* carve_rows(const Cfg& c, float* base, Rows& r, bool bwd) {
  r.xs = base;
  r.xo = r.xs + c.ldx;
  r.prj = r.xo + c.ldx;
  float* q = r.prj + (bwd ? 2 : 1) * c.ldy;
  r.so = q;
  q += c.SP;
  r.lt = q;
  q += c.SP;
  r.h = q;
  q += c.SP;
  r.p = q;
  q += c.SP;
  r.d = q;
  q += c.SP;
  r.mean = q;
  q += c.SP;
  r.sd = q;
  q += c.SP;
  return q;
}
__device__ __forceinline__ void stage_rows(const Cfg& c, const PeriodArgs& a, const Rows& r, const float* __restrict__ x,
                                           const float* __restrict__ prj, int b, int lane) {
  for (int k = lane * 4; k < c.ldx; k += 128)
    *reinterpret_cast<float4*>(r.xs + k) = *reinterpret_cast<const float4*>(x + k);
  for (int k = lane * 4; k < c.ldy; k += 128)
    *reinterpret_cast<float4*>(r.prj + k) = *reinterpret_cast<const float4*>(prj + k);
  const size_t o = static_cast<size_t>(b) * c.S;
  for (int s = lane; s < c.SP; s += 32) {
    const bool v = s < c.S;
    r.lt[s] = v ? __ldg(a.st.lead_times + o + s) : 1.f;
    r.h[s] = v ? __ldg(a.st.holding_costs + o + s) : 0.f;
    r.p[s] = v ? __ldg(a.st.underage_costs + o + s) : 0.f;
    r.mean[s] = v ? __ldg(a.st.mean + o + s) : 0.f;
    r.sd[s] = v ? __ldg(a.st.std + o + s) : 0.f;
    r.d[s] = v ? demand_of(c, a, b, s) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward head
// ------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kMaxWarps * 32)
sym_head_fwd_kernel(Cfg c, PeriodArgs a, const float* __restrict__ params, const float* __restrict__ X,
                    const float* __restrict__ PRJ, float* __restrict__ Xn, float* __restrict__ Xn_hi,
                    float* __restrict__ Xn_lo, float* __restrict__ so_tape, float* __restrict__ cost_b,
                    float* __restrict__ report_b, float* __restrict__ reward_t, int per_warp) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights(c, params, Ws, false);  // parameters do not depend on the predecessor kernel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  Rows r;
  float* q = carve_rows(c, smem + weight_floats(c, false) + static_cast<size_t>(warp) * per_warp, r, false);
  float* loc = q;
  q += NS * 32 * c.s_xs;
  float* hid = q;
  q += NS * 32 * HS;
  float* hw = q;
  pdl_wait();
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = gwarp; b < c.Bp; b += nwarps) {
    float* xn_row = Xn + static_cast<size_t>(b) * c.ldx;
    if (b >= c.B) {  // tile-padding rows stay exactly zero
      for (int k = lane * 4; k < c.ldx; k += 128) {
        *reinterpret_cast<float4*>(xn_row + k) = zero4;
        if (Xn_hi) {
          *reinterpret_cast<float4*>(Xn_hi + static_cast<size_t>(b) * c.ldx + k) = zero4;
          *reinterpret_cast<float4*>(Xn_lo + static_cast<size_t>(b) * c.ldx + k) = zero4;
        }
      }
      continue;
    }
    __syncwarp();
    stage_rows(c, a, r, X + static_cast<size_t>(b) * c.ldx, PRJ + static_cast<size_t>(b) * c.ldy, b, lane);
    float wh_hold = __ldg(a.st.warehouse_holding_costs + b);
    float wh_lead = __ldg(a.st.warehouse_lead_times + b);
    float wh_edge = c.has_edge ? __ldg(a.st.warehouse_edge_costs + b) : 0.f;
    __syncwarp();
    const float* x = r.xs;
    // ---- store net for every store: lane = store, NS * 32 stores per pass
    for (int s0 = 0; s0 < c.S; s0 += NS * 32) {
      const float* locr[NS];
      float* hidr[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = s0 + j * 32 + lane;
        float* row = loc + (j * 32 + lane) * c.s_xs;
        build_loc_row(c, row, x, s < c.S ? s : c.S - 1, r.mean, r.sd, r.p, r.lt);
        locr[j] = row;
        hidr[j] = hid + (j * 32 + lane) * HS;
      }
      float y[NS];
      store_net_fwd<NS>(c, Ws, r.prj, locr, hidr, 0, y, c.tc_fwd != 0);
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int s = s0 + j * 32 + lane;
        if (s < c.SP) r.so[s] = s < c.S ? act_fwd(c.s_oact, y[j]) : 0.f;
      }
    }
    __syncwarp();
    // ---- warehouse net (lane = unit)
    const float yw = wh_net_fwd(c, Ws, r.prj + H, x + c.nS, hw, lane);
    float aw = act_fwd(c.w_oact, yw) * c.wub;
    if (c.discrete) aw = rintf(aw);
    // ---- proportional allocation: stores get min(1, on-hand / (sum of requests + eps)) of what they ask for
    if (so_tape) {
      float* so_row = so_tape + static_cast<size_t>(b) * c.ldo;
      for (int k = lane * 4; k < c.ldo; k += 128)
        *reinterpret_cast<float4*>(so_row + k) = *reinterpret_cast<const float4*>(r.so + k);
    }
    float part = 0.f;
    for (int s = lane; s < c.S; s += 32) part += r.so[s];
    const float tot = warp_sum(part);
    const float W0 = x[c.nS];
    const float scale = fminf(W0 / (tot + c.eps), 1.f);
    __syncwarp();
    // ---- stores
    float* xn = r.xo;
    float cost = 0.f, drawn = 0.f;
    for (int s = lane; s < c.S; s += 32) {
      float al = r.so[s] * scale;
      if (c.discrete) al = rintf(al);
      drawn += al;
      const float* xs = x + s * c.L;
      float* xo = xn + s * c.L;
      const float on_hand = xs[0];
      const float d = r.d[s];
      const float raw = on_hand - d;
      const float h = r.h[s], p = r.p[s];
      cost += c.profit ? (-p * fminf(on_hand, d) + h * relu0(raw)) : (p * relu0(-raw) + h * relu0(raw));
      const float post = c.lost ? relu0(raw) : raw;
      xo[0] = post + xs[1];
      for (int k = 1; k < c.L - 1; ++k) xo[k] = xs[k + 1];
      xo[c.L - 1] = 0.f;
      if (al != 0.f) {
        const int slot = static_cast<int>(r.lt[s]) - 1;
        if (slot >= 0 && slot < c.L) xo[slot] += al;
      }
    }
    drawn = warp_sum(drawn);
    // ---- warehouse (lane 0)
    if (lane == 0) {
      const float* xw = x + c.nS;
      float* xo = xn + c.nS;
      const float raw = xw[0] - drawn;
      float cw = wh_hold * relu0(raw);
      if (c.has_edge) cw += wh_edge * aw;
      cost += cw;
      xo[0] = raw + xw[1];
      for (int k = 1; k < c.Lw - 1; ++k) xo[k] = xw[k + 1];
      xo[c.Lw - 1] = 0.f;
      if (aw != 0.f) {
        const int slot = static_cast<int>(wh_lead) - 1;
        if (slot >= 0 && slot < c.Lw) xo[slot] += aw;
      }
    }
    for (int k = c.nS + c.Lw + lane; k < c.ldx; k += 32) xn[k] = 0.f;
    __syncwarp();
    store_row_split(xn_row, Xn_hi ? Xn_hi + static_cast<size_t>(b) * c.ldx : nullptr,
                    Xn_lo ? Xn_lo + static_cast<size_t>(b) * c.ldx : nullptr, xn, c.ldx, lane);
    cost = warp_sum(cost);
    if (lane == 0) {
      cost_b[b] += cost;
      if (report_b && a.in_report) report_b[b] += cost;
      if (reward_t) reward_t[b] = cost;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// adjoint head
// ------------------------------------------------------------------------------------------------------------
// TC: the store net's GEMM-shaped steps (recompute, hidden-layer dgrad, weight gradients) on mma.sync (mma32.cuh); the
// gradient accumulators are then mma C fragments instead of the SIMT 4 x KQ tiles.
template <int KQ0, int NHH, bool TC>
__global__ void __launch_bounds__(kMaxWarps * 32)
sym_head_bwd_kernel(Cfg c, PeriodArgs a, const float* __restrict__ params, const float* __restrict__ X,
                    const float* __restrict__ PRJ, const float* __restrict__ so_tape, float* __restrict__ gX,
                    float* __restrict__ gPRJ, float* __restrict__ gPRJ_lo, float rb, float* __restrict__ slabs, int first,
                    int per_warp) {
  HDPO_DYN_SMEM(float, smem);
  float* Ws = smem;
  stage_weights(c, params, Ws, true);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int gwarp = blockIdx.x * wpc + warp, nwarps = gridDim.x * wpc;
  Rows r;
  float* q = carve_rows(c, smem + weight_floats(c, true) + static_cast<size_t>(warp) * per_warp, r, true);
  float* g = r.xo;               // adjoint row: staged, updated in place, written back
  float* gprj = r.prj + c.ldy;   // adjoint of the projection row
  float* ga = q;                 // adjoint of the allocations, then of the store outputs
  q += c.SP;
  float* loc = q;                // [32][s_xs] local input rows of the current pass
  q += 32 * c.s_xs;
  float* Hb = q;                 // [NHH + 1][32][HS] activations -> overwritten by the pre-activation adjoints
  q += (NHH + 1) * 32 * HS;
  float* Gy = q;                 // [32] output adjoints of the current pass
  q += 32;
  float* hw = q;                 // warehouse net activations [w_nhh + 1][H]
  q += (kMaxHH + 1) * H;
  float* gzw = q;                // warehouse net: broadcast row of pre-activation adjoints (+ spare row)
  q += 2 * H;
  float* wacc = q;               // warehouse net gradient accumulators
  constexpr int HL = 32 * HS;
  const int lws = c.Lw | 1;
  float* wa_w0 = wacc;                      // [H][lws]
  float* wa_wh = wa_w0 + H * lws;           // [w_nhh][H][WS]
  float* wa_bh = wa_wh + c.w_nhh * H * WS;  // [w_nhh][H]
  float* wa_wo = wa_bh + c.w_nhh * H;       // [H]
  float* wa_bo = wa_wo + H;                 // [1]
  for (int i = lane; i < wacc_floats(c); i += 32) wacc[i] = 0.f;

  // register-resident gradient tiles of the store net
  small::WgradAcc<TC ? 1 : KQ0> a0;
  small::WgradAcc<TC ? 1 : 8> ah[NHH > 0 ? NHH : 1];
  float ao = 0.f, bo = 0.f;  // lane = k for ao
  a0.clear();
#pragma unroll
  for (int l = 0; l < (NHH > 0 ? NHH : 1); ++l) ah[l].clear();
  // tensor-core mode: mma C fragments (element i of [mt][nt]: n = 16 mt + (lane >> 2) + 8 (i >> 1),
  // k = 8 nt + 2 (lane & 3) + (i & 1)) and the bias sums with lane = n
  constexpr int NT0 = KQ0 / 2;  // 8-column blocks of the local first-layer inputs (s_in4 = 4 KQ0)
  float a0f[2][TC ? NT0 : 1][4], ahf[NHH > 0 ? NHH : 1][2][TC ? 4 : 1][4], bhf[NHH > 0 ? NHH : 1];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int nt = 0; nt < (TC ? NT0 : 1); ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) a0f[mt][nt][i] = 0.f;
#pragma unroll
    for (int l = 0; l < (NHH > 0 ? NHH : 1); ++l) {
      bhf[l] = 0.f;
#pragma unroll
      for (int nt = 0; nt < (TC ? 4 : 1); ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) ahf[l][mt][nt][i] = 0.f;
    }
  }

  pdl_wait();
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = gwarp; b < c.Bp; b += nwarps) {
    if (b >= c.B) {
      for (int k = lane * 4; k < c.ldy; k += 128) {
        *reinterpret_cast<float4*>(gPRJ + static_cast<size_t>(b) * c.ldy + k) = zero4;
        if (gPRJ_lo) *reinterpret_cast<float4*>(gPRJ_lo + static_cast<size_t>(b) * c.ldy + k) = zero4;
      }
      continue;
    }
    __syncwarp();
    stage_rows(c, a, r, X + static_cast<size_t>(b) * c.ldx, PRJ + static_cast<size_t>(b) * c.ldy, b, lane);
    {
      const float* so_row = so_tape + static_cast<size_t>(b) * c.ldo;
      for (int k = lane * 4; k < c.ldo; k += 128)
        *reinterpret_cast<float4*>(r.so + k) = *reinterpret_cast<const float4*>(so_row + k);
    }
    float* gx_row = gX + static_cast<size_t>(b) * c.ldx;
    for (int k = lane * 4; k < c.ldx; k += 128)
      *reinterpret_cast<float4*>(g + k) = *reinterpret_cast<const float4*>(gx_row + k);
    const float wh_hold = __ldg(a.st.warehouse_holding_costs + b);
    const float wh_lead = __ldg(a.st.warehouse_lead_times + b);
    const float wh_edge = c.has_edge ? __ldg(a.st.warehouse_edge_costs + b) : 0.f;
    __syncwarp();
    const float* x = r.xs;
    // ---- recompute: proportional allocation scale, warehouse net
    float part = 0.f;
    for (int s = lane; s < c.S; s += 32) part += r.so[s];
    const float tot = warp_sum(part);
    const float W0 = x[c.nS];
    const float den = tot + c.eps;
    const float ratio = W0 / den;
    const float scale = fminf(ratio, 1.f);
    const float yw = wh_net_fwd(c, Ws, r.prj + H, x + c.nS, hw, lane);
    const float ow = act_fwd(c.w_oact, yw);
    const float aw = ow * c.wub;
    part = 0.f;
    for (int s = lane; s < c.S; s += 32) part += r.so[s] * scale;
    const float drawn = warp_sum(part);
    // ---- warehouse period adjoint (same scalars in every lane; lane 0 updates the row)
    float gaw = 0.f, g_raw_w;
    {
      float* gw = g + c.nS;
      const float raw = W0 - drawn;
      if (aw != 0.f) {
        const int slot = static_cast<int>(wh_lead) - 1;
        if (slot >= 0 && slot < c.Lw) gaw = gw[slot];
      }
      if (c.has_edge) gaw += rb * wh_edge;
      const float gn0 = gw[0];
      g_raw_w = rb * wh_hold * ge0(raw) + gn0;
      __syncwarp();
      if (lane == 0) {
        for (int k = c.Lw - 1; k >= 2; --k) gw[k] = gw[k - 1];
        gw[1] = gn0;
        gw[0] = g_raw_w;
      }
    }
    // ---- store period adjoint (lane = store) and the allocation adjoints
    float gsc = 0.f;
    for (int s = lane; s < c.SP; s += 32) {
      float gal = 0.f;
      if (s < c.S) {
        const float* xs = x + s * c.L;
        float* gs = g + s * c.L;
        const float on_hand = xs[0];
        const float d = r.d[s];
        const float raw = on_hand - d;
        const float h = r.h[s], p = r.p[s];
        const float al = r.so[s] * scale;
        if (al != 0.f) {
          const int slot = static_cast<int>(r.lt[s]) - 1;
          if (slot >= 0 && slot < c.L) gal = gs[slot];
        }
        gal -= g_raw_w;
        float g0;
        if (c.profit) {
          const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
          g0 = rb * (-p * tie + h * ge0(raw));
        } else {
          g0 = rb * (-p * le0(raw) + h * ge0(raw));
        }
        const float gn0 = gs[0];
        g0 += c.lost ? gn0 * ge0(raw) : gn0;
        for (int k = c.L - 1; k >= 2; --k) gs[k] = gs[k - 1];
        gs[1] = gn0;
        gs[0] = g0;
        gsc += gal * r.so[s];
      }
      ga[s] = gal;
    }
    // ---- proportional allocation adjoint: alloc = so * min(ratio, 1), ratio = W0 / (sum so + eps)
    const float g_scale = warp_sum(gsc);
    const float g_ratio = ratio <= 1.f ? g_scale : 0.f;  // clip(max=1) passes the gradient at the boundary
    const float g_tot = -g_ratio * W0 / (den * den);
    __syncwarp();
    if (lane == 0) g[c.nS] += g_ratio / den;
    for (int s = lane; s < c.SP; s += 32) ga[s] = s < c.S ? ga[s] * scale + g_tot : 0.f;
    // ---- warehouse net adjoint (lane = unit)
    float gprj_w;
    {
      const float gyw = gaw * c.wub * act_grad(c.w_oact, yw, ow);
      const float* hl = hw + c.w_nhh * H;
      wa_wo[lane] += gyw * hl[lane];
      if (lane == 0) wa_bo[0] += gyw;
      float gz = Ws[c.m_w_wo + lane] * gyw * act_grad_from_out(c.w_hact, hl[lane]);
      for (int l = c.w_nhh - 1; l >= 0; --l) {
        __syncwarp();
        gzw[lane] = gz;
        __syncwarp();
        float* acc = wa_wh + l * H * WS + lane * WS;
        const float* hin = hw + l * H;
#pragma unroll 8
        for (int k = 0; k < H; ++k) acc[k] = fmaf(gz, hin[k], acc[k]);
        wa_bh[l * H + lane] += gz;
        const float* Wt = Ws + c.m_w_wth[l] + lane * WS;  // row k = lane: W[n][lane] over n
        float gh = 0.f;
#pragma unroll 8
        for (int n = 0; n < H; ++n) gh = fmaf(Wt[n], gzw[n], gh);
        gz = gh * act_grad_from_out(c.w_hact, hin[lane]);
      }
      // layer 0: local weights, projection adjoint, pipeline adjoint
      const float* xw = x + c.nS;
      for (int k = 0; k < c.Lw; ++k) wa_w0[lane * lws + k] = fmaf(gz, xw[k], wa_w0[lane * lws + k]);
      gprj_w = gz;
      __syncwarp();
      for (int k = 0; k < c.Lw; ++k) {
        const float gk = warp_sum(Ws[c.m_w_wt0 + k * WS + lane] * gz);
        if (lane == 0) g[c.nS + k] += gk;
      }
    }
    __syncwarp();
    // ---- store net adjoint, 32 stores per pass
    float gprj_s = 0.f;  // lane = unit n: sum over the stores of the layer-0 pre-activation adjoints
    float* xrow = loc + lane * c.s_xs;
    float* hrow = Hb + lane * HS;
    for (int s0 = 0; s0 < c.S; s0 += 32) {
      const int s = s0 + lane;
      const bool valid = s < c.S;
      build_loc_row(c, xrow, x, valid ? s : c.S - 1, r.mean, r.sd, r.p, r.lt);
      float y[1];
      {
        const float* xin[1] = {xrow};
        float* hr[1] = {hrow};
        store_net_fwd<1>(c, Ws, r.prj, xin, hr, HL, y, c.tc_recompute != 0);
      }
      const float gy = valid ? ga[s] * act_grad(c.s_oact, y[0], act_fwd(c.s_oact, y[0])) : 0.f;
      Gy[lane] = gy;
      __syncwarp();
      // output layer: dwo[k] += sum_rows gy * h_last[k] (lane = k), dbo
      {
        const float* Hl = Hb + NHH * HL;
        for (int cidx = 0; cidx < 32; ++cidx) {
          const float gv = Gy[cidx];
          ao = fmaf(gv, Hl[cidx * HS + lane], ao);
          bo += gv;
        }
      }
      __syncwarp();
      // thread-local: gz = wo * gy * act'(h_last), written over the activation row
      float2 gz[H / 2];
      {
        float* hl = hrow + NHH * HL;
        dispatch_act(c.s_hact, [&](auto tag) {
          constexpr int ACT = decltype(tag)::value;
#pragma unroll
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 wv = reinterpret_cast<const float4*>(Ws + c.m_s_wo)[k4];
            const float4 hv = reinterpret_cast<const float4*>(hl)[k4];
            gz[2 * k4 + 0] = make_float2(wv.x * gy * act_grad_out_t<ACT>(hv.x), wv.y * gy * act_grad_out_t<ACT>(hv.y));
            gz[2 * k4 + 1] = make_float2(wv.z * gy * act_grad_out_t<ACT>(hv.z), wv.w * gy * act_grad_out_t<ACT>(hv.w));
            reinterpret_cast<float4*>(hl)[k4] =
                make_float4(gz[2 * k4].x, gz[2 * k4].y, gz[2 * k4 + 1].x, gz[2 * k4 + 1].y);
          }
        });
      }
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
          dispatch_act(c.s_hact, [&](auto tag) {
            constexpr int ACT = decltype(tag)::value;
            mma32::dgrad_inplace(Ws + c.m_s_whk_hi[l], Ws + c.m_s_whk_lo[l], HS, Gz, HS, Hb + l * HL, HS, lane,
                                 [](float y) { return act_grad_out_t<ACT>(y); });
          });
          float* hq = hrow + l * HL;  // this lane's row of the new pre-activation adjoints
#pragma unroll
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 v = reinterpret_cast<const float4*>(hq)[k4];
            gz[2 * k4 + 0] = make_float2(v.x, v.y);
            gz[2 * k4 + 1] = make_float2(v.z, v.w);
          }
#endif
        } else {
        small::wgrad_tile<8>(Hb + (l + 1) * HL, HS, Hb + l * HL, HS, lane, ah[l]);
        __syncwarp();
        float* hp = hrow + l * HL;
        const float* Wt = Ws + c.m_s_wth[l];
        dispatch_act(c.s_hact, [&](auto tag) {
          constexpr int ACT = decltype(tag)::value;
#pragma unroll 1
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 hv = reinterpret_cast<const float4*>(hp)[k4];
            float4 rr;
            rr.x = small::dgrad_dot(Wt + (4 * k4 + 0) * H, gz) * act_grad_out_t<ACT>(hv.x);
            rr.y = small::dgrad_dot(Wt + (4 * k4 + 1) * H, gz) * act_grad_out_t<ACT>(hv.y);
            rr.z = small::dgrad_dot(Wt + (4 * k4 + 2) * H, gz) * act_grad_out_t<ACT>(hv.z);
            rr.w = small::dgrad_dot(Wt + (4 * k4 + 3) * H, gz) * act_grad_out_t<ACT>(hv.w);
            reinterpret_cast<float4*>(hp)[k4] = rr;
          }
        });
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 v = reinterpret_cast<const float4*>(hp)[k4];
          gz[2 * k4 + 0] = make_float2(v.x, v.y);
          gz[2 * k4 + 1] = make_float2(v.z, v.w);
        }
        }  // SIMT form
      }
      // layer 0: local weight gradient, projection adjoint (column sums over the rows), pipeline adjoint
      __syncwarp();
      if constexpr (TC) {
#ifndef HDPO_EMU
        mma32::wgrad<NT0>(Hb, HS, loc, c.s_xs, lane, a0f);
#endif
      } else {
        small::wgrad_tile<KQ0>(Hb, HS, loc, c.s_xs, lane, a0);
      }
      for (int cidx = 0; cidx < 32; ++cidx) gprj_s += Hb[cidx * HS + lane];
      if (valid) {
        float* gs = g + s * c.L;
        const float* Wt0 = Ws + c.m_s_wt0;
        for (int k = 0; k < c.L; ++k) gs[k] += small::dgrad_dot(Wt0 + k * H, gz);
      }
      __syncwarp();
    }
    // ---- write back: adjoint row of X_t (direct part), adjoint of the projection row
    gprj[lane] = gprj_s;
    gprj[H + lane] = gprj_w;
    __syncwarp();
    for (int k = lane * 4; k < c.ldx; k += 128)
      *reinterpret_cast<float4*>(gx_row + k) = *reinterpret_cast<const float4*>(g + k);
    if (gPRJ_lo)
      store_row_split(nullptr, gPRJ + static_cast<size_t>(b) * c.ldy, gPRJ_lo + static_cast<size_t>(b) * c.ldy, gprj,
                      c.ldy, lane);
    else
      store_row_split(gPRJ + static_cast<size_t>(b) * c.ldy, nullptr, nullptr, gprj, c.ldy, lane);
  }
  __syncwarp();

  // ---- add this launch's gradients to the warp's slab (first launch of a sweep overwrites)
  float* out = slabs + static_cast<size_t>(gwarp) * c.q_total;
  auto put = [&](int at, float v) { out[at] = first ? v : out[at] + v; };
  const int ni = lane >> 2, ki = lane & 3;
  if constexpr (TC) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = 16 * mt + ni + 8 * (i >> 1), k = 2 * ki + (i & 1);
#pragma unroll
        for (int nt = 0; nt < (TC ? NT0 : 1); ++nt) put(c.q_s_w0 + n * c.s_in4 + 8 * nt + k, a0f[mt][nt][i]);
#pragma unroll
        for (int l = 0; l < NHH; ++l)
#pragma unroll
          for (int nt = 0; nt < (TC ? 4 : 1); ++nt) put(c.q_s_wh[l] + n * H + 8 * nt + k, ahf[l][mt][nt][i]);
      }
#pragma unroll
    for (int l = 0; l < NHH; ++l) put(c.q_s_bh[l] + lane, bhf[l]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = 4 * ni + i;
#pragma unroll
      for (int qq = 0; qq < (TC ? 1 : KQ0); ++qq) put(c.q_s_w0 + n * c.s_in4 + ki * KQ0 + qq, a0.at(i, qq));
    }
#pragma unroll
    for (int l = 0; l < NHH; ++l) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = 4 * ni + i;
#pragma unroll
        for (int qq = 0; qq < (TC ? 1 : 8); ++qq) put(c.q_s_wh[l] + n * H + ki * 8 + qq, ah[l].at(i, qq));
        if (ki == 0) put(c.q_s_bh[l] + n, ah[l].bias[i]);
      }
    }
  }
  put(c.q_s_wo + lane, ao);
  if (lane == 0) put(c.q_s_bo, bo);
  {
    const int lw4 = (c.Lw + 3) & ~3;
    for (int k = 0; k < c.Lw; ++k) put(c.q_w_w0 + lane * lw4 + k, wa_w0[lane * lws + k]);
    for (int l = 0; l < c.w_nhh; ++l) {
      for (int k = 0; k < H; ++k) put(c.q_w_wh[l] + lane * H + k, wa_wh[l * H * WS + lane * WS + k]);
      put(c.q_w_bh[l] + lane, wa_bh[l * H + lane]);
    }
    put(c.q_w_wo + lane, wa_wo[lane]);
    if (lane == 0) put(c.q_w_bo, wa_bo[0]);
  }
}

// grad[block element] = sum over the slabs in fixed order; one CTA row per parameter block
__global__ void __launch_bounds__(256) sym_reduce_kernel(Cfg c, const float* __restrict__ slabs, int n_slabs,
                                                         float* __restrict__ grad) {
  pdl_wait();
  const Block bk = c.blk[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bk.rows * bk.cols) return;
  const int rr = i / bk.cols, cc = i % bk.cols;
  const float* src = slabs + bk.q + rr * bk.q_ld + cc;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int w = 0;
  for (; w + 3 < n_slabs; w += 4) {
    s0 += src[static_cast<size_t>(w) * c.q_total];
    s1 += src[static_cast<size_t>(w + 1) * c.q_total];
    s2 += src[static_cast<size_t>(w + 2) * c.q_total];
    s3 += src[static_cast<size_t>(w + 3) * c.q_total];
  }
  for (; w < n_slabs; ++w) s0 += src[static_cast<size_t>(w) * c.q_total];
  grad[bk.dst + rr * bk.dst_ld + cc] = (s0 + s1) + (s2 + s3);
}

// projection layer of the trunk: packed row n < 32 -> store net row n (context columns), 32 <= n -> warehouse net
__device__ __forceinline__ float tf32_round_pack(float x) { return tf32_round(x); }
__global__ void __launch_bounds__(256) sym_pack_projection_kernel(Cfg c, const float* __restrict__ params, int Kp,
                                                                  float* __restrict__ Wp, float* __restrict__ bp,
                                                                  float* __restrict__ W_lo, float* __restrict__ WT,
                                                                  float* __restrict__ WT_lo) {
  pdl_wait();
  const int Np = 2 * H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np * Kp) {
    const int n = i / Kp, k = i % Kp;
    float w = 0.f;
    if (k < c.C) {
      if (n < H) {
        if (n < c.s_w[0]) w = params[c.g_s_w0 + n * c.s_ld0 + c.s_in + k];
      } else if (n - H < c.w_w[0]) {
        w = params[c.g_w_w0 + (n - H) * c.w_ld0 + c.Lw + k];
      }
    }
    if (W_lo) {
      const float hi = tf32_round_pack(w), lo = tf32_round_pack(w - hi);
      Wp[i] = hi;
      W_lo[i] = lo;
      WT[static_cast<size_t>(k) * Np + n] = hi;
      WT_lo[static_cast<size_t>(k) * Np + n] = lo;
    } else {
      Wp[i] = w;
    }
  }
  if (i < Np) {
    float bv = 0.f;
    if (i < H) {
      if (i < c.s_w[0]) bv = params[c.g_s_b0 + i];
    } else if (i - H < c.w_w[0]) {
      bv = params[c.g_w_b0 + (i - H)];
    }
    bp[i] = bv;
  }
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------
int pack_projection(const Cfg& c, const float* params, int Kp, float* Wp, float* bp, float* W_lo, float* WT, float* WT_lo,
                    void* stream) {
  auto k = sym_pack_projection_kernel;
  HDPO_LAUNCH_PDL(k, ceil_div(2 * H * Kp, 256), 256, 0, stream, c, params, Kp, Wp, bp, W_lo, WT, WT_lo);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

template <typename K>
static int set_smem(K k, size_t bytes) {
#ifndef HDPO_EMU
  if (bytes > 48 * 1024)
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemMax)));
#else
  (void)k;
  (void)bytes;
#endif
  return HDPO_OK;
}

int head_fwd(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ, float* Xn,
             float* Xn_hi, float* Xn_lo, float* so_tape, float* cost_b, float* report_b, float* reward_t, void* stream) {
  const int per_warp = fwd_smem_floats_per_warp(c);
  int wpc = warps_per_cta(c, per_warp, c.Bp, false);
  if (wpc > 4 && !getenv("HDPO_SYM_WPC")) wpc = 4;  // measured: 2 resident CTAs of 4 warps beat 1 of 8 (finer tail)
  const size_t smem = smem_bytes(c, wpc, per_warp, false);
  int grid = ceil_div(c.Bp, wpc);
  const int cap = max_resident_ctas(smem);
  if (grid > cap) grid = cap;
  HDPO_REQUIRE(smem <= kSmemMax, "symmetry-aware head: %zu bytes of shared memory needed", smem);
  int rc;
  if (fwd_ns(c) == 2) {
    auto k = sym_head_fwd_kernel<2>;
    if ((rc = set_smem(k, smem))) return rc;
    HDPO_LAUNCH_PDL(k, grid, wpc * 32, smem, stream, c, a, params, X, PRJ, Xn, Xn_hi, Xn_lo, so_tape, cost_b, report_b,
                    reward_t, per_warp);
  } else {
    auto k = sym_head_fwd_kernel<1>;
    if ((rc = set_smem(k, smem))) return rc;
    HDPO_LAUNCH_PDL(k, grid, wpc * 32, smem, stream, c, a, params, X, PRJ, Xn, Xn_hi, Xn_lo, so_tape, cost_b, report_b,
                    reward_t, per_warp);
  }
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

template <int KQ0, int NHH, bool TC>
static int launch_bwd_tc(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ,
                      const float* so_tape, float* gX, float* gPRJ, float* gPRJ_lo, float rb, float* slabs, int first,
                      void* stream) {
  int grid, wpc;
  bwd_shape(c, &grid, &wpc);
  const int per_warp = bwd_smem_floats_per_warp(c);
  const size_t smem = smem_bytes(c, wpc, per_warp, true);
  HDPO_REQUIRE(smem <= kSmemMax, "symmetry-aware adjoint head: %zu bytes of shared memory needed", smem);
  auto k = sym_head_bwd_kernel<KQ0, NHH, TC>;
  int rc = set_smem(k, smem);
  if (rc) return rc;
  HDPO_LAUNCH_PDL(k, grid, wpc * 32, smem, stream, c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first,
                  per_warp);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

template <int KQ0, int NHH>
static int launch_bwd(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ,
                      const float* so_tape, float* gX, float* gPRJ, float* gPRJ_lo, float rb, float* slabs, int first,
                      void* stream) {
#ifndef HDPO_EMU
  if (c.tc) return launch_bwd_tc<KQ0, NHH, true>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
#endif
  return launch_bwd_tc<KQ0, NHH, false>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
}

template <int KQ0>
static int launch_bwd_nhh(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ,
                          const float* so_tape, float* gX, float* gPRJ, float* gPRJ_lo, float rb, float* slabs, int first,
                          void* stream) {
  switch (c.s_nhh) {
    case 0: return launch_bwd<KQ0, 0>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
    case 1: return launch_bwd<KQ0, 1>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
    case 2: return launch_bwd<KQ0, 2>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
  }
  set_error("unsupported store-net depth %d", c.s_nhh);
  return HDPO_E_INVALID;
}

int head_bwd(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ,
             const float* so_tape, float* gX, float* gPRJ, float* gPRJ_lo, float rb, float* slabs, int first,
             void* stream) {
  switch (c.s_kq) {
    case 2: return launch_bwd_nhh<2>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
    case 4: return launch_bwd_nhh<4>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
    case 8: return launch_bwd_nhh<8>(c, a, params, X, PRJ, so_tape, gX, gPRJ, gPRJ_lo, rb, slabs, first, stream);
  }
  set_error("unsupported store-net input tile %d", c.s_kq);
  return HDPO_E_INVALID;
}

int reduce_slabs(const Cfg& c, const float* slabs, float* grad, void* stream) {
  int max_elems = 1;
  for (int i = 0; i < c.n_blocks; ++i)
    if (c.blk[i].rows * c.blk[i].cols > max_elems) max_elems = c.blk[i].rows * c.blk[i].cols;
  auto k = sym_reduce_kernel;
  HDPO_LAUNCH_PDL(k, dim3(ceil_div(max_elems, 256), c.n_blocks), 256, 0, stream, c, slabs, bwd_warps(c), grad);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

}  // namespace sym
}  // namespace hdpo
