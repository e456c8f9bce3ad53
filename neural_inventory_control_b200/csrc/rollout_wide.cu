// rollout_wide.cu - K1/K2 for wide policy nets over many stores (VanillaWarehouse: one or many warehouses).
//
// Replaces trainer.py:181-216 + neural_networks.py:369-427 + environment.py:110-270 and their autograd for a whole
// batch. Unlike the small-net path the policy MLP here is a real GEMM (153->512->512->512->51 per scenario), so the
// rollout is a pipeline of tile kernels per period instead of one persistent kernel:
//
//   forward, period t :  X_t --[gemm+bias+act] x n_layers--> Y_t --[warehouse_head_fwd: masked softmax x on-hand,
//                        sigmoid x bound, store / warehouse dynamics, cost]--> X_{t+1}
//   adjoint, period t :  gX_{t+1} --[warehouse_head_bwd]--> gY_t, gX_t(direct) --[dgrad gemm (x act') ...]--> gX_t
//   after the sweep   :  dW_l = sum_t gz_{l,t}^T h_{l-1,t} as ONE split-K GEMM per layer over all T*B rows
//
// HBM layout (180 GB part): everything the adjoint needs is SAVED, not recomputed - the state tape X[T+1][Bp][w0],
// every layer output ACT_l[T][Bp][w_{l+1}] and every pre-activation adjoint GZ_l[T][Bp][w_{l+1}] (cfg 4 at
// B = 8192: ~5.6 GB). Rows are scenarios (padded to 128), columns padded to 64, so all tiles are full and all
// float4 accesses aligned; weights are re-packed into zero-padded [N][K] slabs once per call.
//
// This file holds the fp32 SIMT tile GEMM (parity mode). The tcgen05 3xTF32 GEMM replaces `sgemm` call sites
// one-for-one (same operand layouts) - see gemm_tc.cuh.
#include "rollout_wide.cuh"
#include "rollout_sym.cuh"
#include "gemm_tc.cuh"
#include "wide_persist.cuh"

#include <cstdlib>
#ifndef HDPO_EMU
#include <mutex>
#endif

namespace hdpo {
namespace wide {

constexpr int kRowPad = 128;   // scenarios padded to the GEMM M tile
constexpr int kColPad = 64;    // layer widths padded to the GEMM N tile
constexpr int kMaxStoresPerWarp = 256;
constexpr int kSplitK = 16;
constexpr size_t kHeadSmemMax = 200 * 1024;  // dynamic shared memory the head kernels may ask for
#ifndef HDPO_HEAD_WARPS
#define HDPO_HEAD_WARPS 4
#endif
constexpr int HEAD_WARPS = HDPO_HEAD_WARPS;  // scenarios (warps) per CTA of the head kernels

// floats of shared memory one warp of a head kernel needs (see HeadSmem)
__host__ __device__ inline int head_smem_floats(int S, int W, int ldx, int ldy, bool bwd) {
  const int SW = S * W;
  const int n = (bwd ? 2 * SW + 64 : SW + 32) + 2 * ldx + (bwd ? 2 : 1) * ldy + SW + 3 * S + 16;
  return (n + 3) & ~3;  // every warp's region starts 16-byte aligned
}


static inline int pad_to(int x, int q) { return (x + q - 1) / q * q; }
static inline size_t a256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

bool supported(const HdpoRolloutDesc* d) {
#ifdef HDPO_EMU
  if (d->precision != HDPO_PREC_FP32) return false;  // tensor cores cannot be emulated
#endif
  if (d->arch == HDPO_ARCH_SYMMETRY_AWARE) return sym::supported(d);
  if (d->arch != HDPO_ARCH_VANILLA_WAREHOUSE) return false;
  const HdpoProblem& pb = d->pb;
  if (pb.W < 1 || pb.E != 0 || pb.W > 8) return false;
  if (pb.S * pb.W > 1024 || pb.S > kMaxStoresPerWarp) return false;
  const HdpoMlp& m = d->master;
  if (m.widths[0] != pb.S * pb.L + pb.W * pb.Lw) return false;
  if (m.widths[m.n_layers] != pb.S * pb.W + pb.W) return false;
  if (m.out_act != HDPO_ACT_NONE) return false;
  const int ldx = pad_to(m.widths[0], kColPad), ldy = pad_to(m.widths[m.n_layers], kColPad);
  if (static_cast<size_t>(HEAD_WARPS) * head_smem_floats(pb.S, pb.W, ldx, ldy, true) * sizeof(float) > kHeadSmemMax)
    return false;
  return true;
}

// ------------------------------------------------------------------------------------------------------------
// workspace plan
// ------------------------------------------------------------------------------------------------------------
struct Plan {
  int n, B, Bp, T, save;
  int sym;                      // SymmetryAware: layers 0..n-2 = context net, layer n-1 = 64-column projection
  int act[HDPO_MAX_LAYERS];     // activation after layer l
  size_t o_so, so_stride;       // sym: store-output tape [T][Bp][ldo]
  size_t o_slab;                // sym: per-warp gradient slabs of the local nets
  int tc, n_pass;  // tensor-core mode: every GEMM operand is kept as a (tf32 hi, remainder lo) pair
  int wg_kps, wg_splits;  // tensor-core weight gradient: contraction rows per partial slice, number of slices
  int w[HDPO_MAX_LAYERS + 1], wp[HDPO_MAX_LAYERS + 1];
  int gw[HDPO_MAX_LAYERS], gb[HDPO_MAX_LAYERS];  // offsets in the flat (state_dict) parameter vector
  int P;
  size_t o_W[HDPO_MAX_LAYERS], o_b[HDPO_MAX_LAYERS];       // packed weights / biases
  size_t o_X, o_act[HDPO_MAX_LAYERS], o_gz[HDPO_MAX_LAYERS];  // tapes
  size_t o_gx, o_part, o_bpart, total;                      // state adjoint, split-K partials
  // split-K of the two THIN GEMMs on the per-period chain (few CTAs, K = the hidden width: their mainloop is a chain of
  // TMA round trips): the output layer writes y_split partial products that the forward head sums (and writes to the Y
  // tape), the first layer's dgrad writes gx_split partials that the next adjoint head adds to its gX row
  int y_split, gx_split;
  size_t o_ypart, o_gxpart;
  int wg_group;                    // > 0: the weight-gradient GEMMs run in groups of this many periods on a second
                                   // (low-priority) stream WHILE the adjoint sweep is still going (wgrad overlap)
  size_t o_part_l[HDPO_MAX_LAYERS];  // partial slices of layer l (= o_part for every layer unless wg_group > 0)
  size_t o_csum[HDPO_MAX_LAYERS];  // tensor-core mode: [T*Bp/32][wp] column sums of 32-row blocks of gz_l (hidden l)
  // tensor-core mode extras: lo halves, transposed weights (dgrad B operand), split state tape
  size_t o_W_lo[HDPO_MAX_LAYERS], o_WT[HDPO_MAX_LAYERS], o_WT_lo[HDPO_MAX_LAYERS];
  size_t o_X_hi, o_X_lo, o_act_lo[HDPO_MAX_LAYERS], o_gz_lo[HDPO_MAX_LAYERS];
  size_t x_stride, act_stride[HDPO_MAX_LAYERS];              // floats per period
  int max_wk;
  int persist;                  // persistent one-launch sweeps (wide_persist.cu): rows padded to 256, one chunk
  size_t o_extra, extra_bytes;  // their private scratch
};

static bool use_persist(const HdpoRolloutDesc* d) {
#ifdef HDPO_EMU
  (void)d;
  return false;
#else
  return wp::enabled() && wp::eligible(d);
#endif
}

// HDPO_WIDE_WG_GROUP = G: periods per weight-gradient group of the overlapped form (0 = all weight gradients after the
// sweep, the round-1 form). The per-period chain leaves a quarter of the SMs idle (dependent 15 us launches, DESIGN
// section 4); the weight gradients are independent of that chain once the gz rows of a period exist.
static int g_wg_group = -1;    // < 0: not read yet (HDPO_WIDE_WG_GROUP, default 5)
static int g_wg_overlap = -2;  // -2: not read yet; -1 = by chunk count, 0 = off, 1 = on (HDPO_WIDE_WG_OVERLAP)
static int wg_group_periods() {
#ifdef HDPO_EMU
  return 0;
#else
  if (g_wg_group < 0) {
    const char* e = getenv("HDPO_WIDE_WG_GROUP");
    g_wg_group = e ? atoi(e) : 5;
    if (g_wg_group < 0) g_wg_group = 0;
  }
  return g_wg_group;
#endif
}
// Measured on B200 (8192 x 50 x 50 stores, 3xTF32): with 3 - 4 concurrent chunk chains the overlapped form changes
// nothing (adjoint 10.8 - 11.1 ms with groups of 2 / 5 / 10 / 25 periods against 10.8 - 11.0 without: the weight-gradient
// CTAs take as many SM slots from the chains as they fill), so it is the default only for ONE chunk (<= 2048 scenarios per
// GPU, e.g. BASELINE cfg 5 at 1024 per GPU: adjoint 4.10 -> 3.84 ms). HDPO_WIDE_WG_OVERLAP = 1 / 0 forces it on / off.
static int wg_overlap_mode() {
#ifdef HDPO_EMU
  return 0;
#else
  if (g_wg_overlap == -2) {
    const char* e = getenv("HDPO_WIDE_WG_OVERLAP");
    g_wg_overlap = e ? (atoi(e) != 0) : -1;
  }
  return g_wg_overlap;
#endif
}

static int multi_min_tiles() {
#ifdef HDPO_EMU
  return 1 << 30;
#else
  return wp::multi_min_tiles();
#endif
}

static int requested_chunks(int B, bool sym);
// Split-K of the thin chain GEMMs: HDPO_WIDE_KSPLIT = 1 / 0 forces it on / off; by default it is on for ONE chunk only.
// Measured on B200 (3xTF32, ms per step off / on): one chunk - one_warehouse 1024 scenarios 6.25 / 5.71, many_warehouses
// 1024 7.22 / 6.84 (the chain is latency-bound, the extra CTAs run on idle SMs); four chunks - one_warehouse 8192
// 16.0 / 17.1, many_warehouses 8192 21.4 / 23.8 (four times the CTAs, each holding an SM's shared memory, compete with
// the other chunk chains).
static int g_ksplit = -2;  // -2: not read yet; -1 = by chunk count, 0 = off, 1 = on
static bool thin_ksplit_enabled(int n_chunks) {
  if (g_ksplit == -2) {
    const char* e = getenv("HDPO_WIDE_KSPLIT");
    g_ksplit = e ? (atoi(e) != 0) : -1;
  }
  return g_ksplit == 1 || (g_ksplit == -1 && n_chunks == 1);
}

static Plan make_plan(const HdpoRolloutDesc* d, int Bc) {
  Plan p;
  p.persist = use_persist(d) ? 1 : 0;
  const HdpoMlp& m = d->master;
  p.sym = d->arch == HDPO_ARCH_SYMMETRY_AWARE;
  p.n = m.n_layers + (p.sym ? 1 : 0);
  p.B = Bc;
  p.Bp = pad_to(p.B > 0 ? p.B : 1, p.persist ? 256 : kRowPad);
  p.T = d->T;
  p.save = d->save_for_backward;
  p.tc = d->precision != HDPO_PREC_FP32;
  p.n_pass = d->precision == HDPO_PREC_TF32 ? 1 : 3;
  int off = 0;
  for (int i = 0; i <= p.n; ++i) {
    p.w[i] = (p.sym && i == p.n) ? 64 : m.widths[i];
    p.wp[i] = pad_to(p.w[i], kColPad);
  }
  for (int l = 0; l < p.n; ++l) {
    if (p.sym)
      p.act[l] = l + 2 < p.n ? m.hidden_act : (l + 2 == p.n ? m.out_act : HDPO_ACT_NONE);
    else
      p.act[l] = l + 1 < p.n ? m.hidden_act : m.out_act;
    p.gw[l] = off;
    off += p.w[l + 1] * p.w[l];
    p.gb[l] = off;
    off += p.w[l + 1];
  }
  if (p.tc && !p.persist) {
    // the multi-tile GEMM works on 256-row tiles: pad to them when any layer of this batch would be routed there
    int widest = 0;
    for (int i = 0; i <= p.n; ++i) widest = p.wp[i] > widest ? p.wp[i] : widest;
    const int bp2 = pad_to(p.B > 0 ? p.B : 1, 256);
    if ((bp2 / 256) * ((widest + 127) / 128) >= multi_min_tiles()) p.Bp = bp2;
  }
  p.P = off;  // sym: overwritten below with the parameter count of all three nets (the projection has no own block)
  sym::Cfg sc{};
  if (p.sym) {
    sym::build_cfg(d, p.B, p.Bp, p.wp[0], p.wp[p.n], &sc);
    p.P = sc.P;
  }
  const size_t f = sizeof(float);
  size_t o = 0;
  auto take = [&](size_t n_floats) {
    size_t at = o;
    o += a256(n_floats * f);
    return at;
  };
  p.max_wk = 0;
  for (int l = 0; l < p.n; ++l) {
    p.o_W[l] = take(static_cast<size_t>(p.wp[l + 1]) * p.wp[l]);
    p.o_b[l] = take(p.wp[l + 1]);
    p.o_W_lo[l] = p.tc ? take(static_cast<size_t>(p.wp[l + 1]) * p.wp[l]) : 0;
    p.o_WT[l] = p.tc ? take(static_cast<size_t>(p.wp[l + 1]) * p.wp[l]) : 0;
    p.o_WT_lo[l] = p.tc ? take(static_cast<size_t>(p.wp[l + 1]) * p.wp[l]) : 0;
    int wk = p.wp[l + 1] * p.wp[l];
    if (wk > p.max_wk) p.max_wk = wk;
  }
  const size_t tslots = p.save ? static_cast<size_t>(p.T) : 1;
  p.x_stride = static_cast<size_t>(p.Bp) * p.wp[0];
  p.o_X = take((p.save ? tslots + 1 : 2) * p.x_stride);
  p.o_X_hi = p.tc ? take(tslots * p.x_stride) : 0;
  p.o_X_lo = p.tc ? take(tslots * p.x_stride) : 0;
  for (int l = 0; l < p.n; ++l) {
    p.act_stride[l] = static_cast<size_t>(p.Bp) * p.wp[l + 1];
    p.o_act[l] = take(tslots * p.act_stride[l]);  // tc: hi half for hidden layers, full fp32 for the output layer
    p.o_act_lo[l] = (p.tc && l + 1 < p.n) ? take(tslots * p.act_stride[l]) : 0;
  }
  for (int l = 0; l < p.n; ++l) {
    p.o_gz[l] = p.save ? take(tslots * p.act_stride[l]) : 0;
    p.o_gz_lo[l] = (p.save && p.tc) ? take(tslots * p.act_stride[l]) : 0;
  }
  p.o_gx = p.save ? take(p.x_stride) : 0;
  for (int l = 0; l < p.n; ++l)
    p.o_csum[l] = (p.save && p.tc && l + 1 < p.n) ? take(tslots * p.act_stride[l] / 32) : 0;
  {
    // short K slices keep the tensor core's truncating accumulation fp32-grade (see gemm_tc.cu)
    const size_t rows = static_cast<size_t>(p.T) * p.Bp;
    // Slice length: the truncation error of a slice grows linearly with its length, the time falls with it (fewer CTA
    // ramp-ups, fewer partial slices to write and reduce). Measured on B200 (8192 x 50 x 50 stores, 3xTF32; rel-L2 distance
    // of the full-batch gradient to the 128-row-slice run | adjoint ms): 512: 6.9e-7 | 10.0; 1024: 1.8e-6 | 9.3; 2048:
    // 4.2e-6 | 9.05; 4096: 8.4e-6. The true-fp32 SIMT path differs from all of them by 6.3e-6 (rounding-level differences of
    // the chaotic 50-store rollouts), so 1024 rows stays well inside the workload's own fp32 floor.
    p.wg_kps = rows % 1024 == 0 ? 1024 : (rows % 512 == 0 ? 512 : (rows % 256 == 0 ? 256 : 128));
    {
      static int kps_env = -1;  // HDPO_WG_KPS: other slice lengths (A/B of accuracy against time, tools/wg_accuracy.py)
      if (kps_env < 0) {
        const char* e = getenv("HDPO_WG_KPS");
        kps_env = e ? atoi(e) : 0;
      }
      if (kps_env > 0 && kps_env % 32 == 0 && rows % static_cast<size_t>(kps_env) == 0) p.wg_kps = kps_env;
    }
    p.wg_splits = static_cast<int>(rows / p.wg_kps);
  }
  p.o_part = p.save ? take(static_cast<size_t>(p.tc ? (p.wg_splits > kSplitK ? p.wg_splits : kSplitK) : kSplitK) * p.max_wk) : 0;
  // overlapped weight gradients: every layer keeps its own partial slices (groups of different layers are in flight
  // at the same time); a group must cover whole K slices
  p.wg_group = 0;
  {
    const int G = wg_group_periods();
    const int mode = wg_overlap_mode();
    const bool wanted = mode == 1 || (mode == -1 && requested_chunks(d->pb.B, p.sym) == 1);
    // HDPO_WIDE_WG_SYM = 0: not for the SymmetryAware trunk (one chunk by design; measured 23.35 -> 23.04 ms per step with it)
    static int sym_ok = -1;
    if (sym_ok < 0) {
      const char* e = getenv("HDPO_WIDE_WG_SYM");
      sym_ok = e ? (atoi(e) != 0) : 1;
    }
    if (wanted && p.save && p.tc && (!p.sym || sym_ok) && !p.persist && G > 0 && G < p.T &&
        (static_cast<size_t>(G) * p.Bp) % static_cast<size_t>(p.wg_kps) == 0 &&
        (p.T % G == 0 || p.Bp % p.wg_kps == 0))  // (the last, shorter group must cover whole slices too)
      p.wg_group = G;
  }
  for (int l = 0; l < p.n; ++l)
    p.o_part_l[l] = p.wg_group ? take(static_cast<size_t>(p.wg_splits > kSplitK ? p.wg_splits : kSplitK) * p.wp[l + 1] * p.wp[l])
                               : p.o_part;  // (layers on the SIMT split-K path write kSplitK slices)
  int max_wp = 0;
  for (int i = 0; i <= p.n; ++i) max_wp = p.wp[i] > max_wp ? p.wp[i] : max_wp;
  p.o_bpart = p.save ? take(static_cast<size_t>(128) * max_wp) : 0;
  p.y_split = p.gx_split = 1;
  p.o_ypart = p.o_gxpart = 0;
#ifndef HDPO_EMU
  if (p.tc && !p.sym && !p.persist && thin_ksplit_enabled(requested_chunks(d->pb.B, p.sym))) {
    static int max_splits = -1;  // HDPO_WIDE_KSPLIT_N: most K slices (default 4)
    if (max_splits < 0) {
      const char* e = getenv("HDPO_WIDE_KSPLIT_N");
      max_splits = e ? atoi(e) : 4;
      if (max_splits < 2) max_splits = 2;
    }
    auto splits = [](int K) {
      if (K < 256 || K % 128 != 0) return 1;
      int n = K / 128 < max_splits ? K / 128 : max_splits;
      while (n > 1 && (K % n != 0 || (K / n) % 32 != 0)) --n;
      return n;
    };
    if (p.act[p.n - 1] == HDPO_ACT_NONE) p.y_split = splits(p.wp[p.n - 1]);  // (an activation cannot act on partials)
    if (p.save) p.gx_split = splits(p.wp[1]);
    if (p.y_split > 1) p.o_ypart = take(static_cast<size_t>(p.y_split) * p.Bp * p.wp[p.n]);
    if (p.gx_split > 1) p.o_gxpart = take(static_cast<size_t>(p.gx_split) * p.Bp * p.wp[0]);
  }
#endif
  p.so_stride = p.sym ? static_cast<size_t>(p.Bp) * sc.ldo : 0;
  p.o_so = (p.sym && p.save) ? take(tslots * p.so_stride) : 0;
  p.o_slab = (p.sym && p.save) ? take(static_cast<size_t>(sym::bwd_warps(sc)) * sc.q_total) : 0;
  p.o_extra = 0;
  p.extra_bytes = 0;
#ifndef HDPO_EMU
  if (p.persist) {
    p.extra_bytes = wp::extra_bytes(d, p.Bp, p.wp, p.n);
    p.o_extra = o;
    o += a256(p.extra_bytes);
  }
#endif
  p.total = o + 256;
  return p;
}


// ------------------------------------------------------------------------------------------------------------
// fp32 SIMT tile GEMM:  C[M,N] = epi( sum_k A(m,k) * B(k,n) ),  all dims multiples of the tile, ld* multiples of 4
//   A_T = false: A(m,k) = A[m*lda + k]      A_T = true: A(m,k) = A[k*lda + m]
//   B_T = false: B(k,n) = B[k*ldb + n]      B_T = true: B(k,n) = B[n*ldb + k]
// ------------------------------------------------------------------------------------------------------------
enum { EPI_BIAS_ACT = 0, EPI_MUL_ACTGRAD = 1, EPI_ACCUM = 2, EPI_SPLITK = 3 };

constexpr int BN = 64, BK = 16, GEMM_THREADS = 256;

struct GemmArgs {
  const float* A;
  const float* B;
  const float* A2;      // optional: operand is A + A2 (hi/lo pairs of the tensor-core mode)
  const float* B2;
  float* C;
  int M, N, K;          // K = contraction length handled by ONE z-slice
  int lda, ldb, ldc;
  const float* bias;    // EPI_BIAS_ACT
  const float* aux;     // EPI_MUL_ACTGRAD: saved layer output, same shape/ld as C
  int act;
  size_t c_slice;       // EPI_SPLITK: floats between z-slices of C
  size_t a_kslice, b_kslice;  // element offset added per z-slice to A / B (k origin of the slice)
};

template <int BM, bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS) sgemm_kernel(GemmArgs g) {
  pdl_wait();
  constexpr int TM = BM / 16;  // rows per thread (8 or 4)
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* __restrict__ A = g.A + blockIdx.z * g.a_kslice;
  const float* __restrict__ Bm = g.B + blockIdx.z * g.b_kslice;
  const float* __restrict__ A2 = g.A2 ? g.A2 + blockIdx.z * g.a_kslice : nullptr;
  const float* __restrict__ B2 = g.B2 ? g.B2 + blockIdx.z * g.b_kslice : nullptr;
  auto ld4 = [](const float* p0, const float* p1, size_t off) {
    float4 v = *reinterpret_cast<const float4*>(p0 + off);
    if (p1) {
      const float4 w = *reinterpret_cast<const float4*>(p1 + off);
      v.x += w.x;
      v.y += w.y;
      v.z += w.z;
      v.w += w.w;
    }
    return v;
  };

  // global -> register staging. A tile: BM x BK floats = BM*4 float4; B tile: 64 x 16 floats = 256 float4.
  constexpr int A_F4 = BM * BK / 4 / GEMM_THREADS;  // 2 (BM=128) or 1 (BM=64)
  float4 ra[A_F4], rb;
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      if (A_T) {  // contiguous along m: idx -> (k = idx / (BM/4), mq = idx % (BM/4))
        const int k = idx / (BM / 4), mq = idx % (BM / 4);
        ra[i] = ld4(A, A2, static_cast<size_t>(k0 + k) * g.lda + m0 + 4 * mq);
      } else {    // contiguous along k: idx -> (m = idx % BM, kq = idx / BM)
        const int m = idx % BM, kq = idx / BM;
        ra[i] = ld4(A, A2, static_cast<size_t>(m0 + m) * g.lda + k0 + 4 * kq);
      }
    }
    if (B_T) {    // B[n*ldb + k], contiguous along k: tid -> (n = tid % 64, kq = tid / 64)
      const int n = tid % BN, kq = tid / BN;
      rb = ld4(Bm, B2, static_cast<size_t>(n0 + n) * g.ldb + k0 + 4 * kq);
    } else {      // B[k*ldb + n], contiguous along n: tid -> (k = tid / 16, nq = tid % 16)
      const int k = tid / (BN / 4), nq = tid % (BN / 4);
      rb = ld4(Bm, B2, static_cast<size_t>(k0 + k) * g.ldb + n0 + 4 * nq);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int idx = tid + i * GEMM_THREADS;
      if (A_T) {
        const int k = idx / (BM / 4), mq = idx % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][k][4 * mq]) = ra[i];
      } else {
        const int m = idx % BM, kq = idx / BM;
        As[buf][4 * kq + 0][m] = ra[i].x;
        As[buf][4 * kq + 1][m] = ra[i].y;
        As[buf][4 * kq + 2][m] = ra[i].z;
        As[buf][4 * kq + 3][m] = ra[i].w;
      }
    }
    if (B_T) {
      const int n = tid % BN, kq = tid / BN;
      Bs[buf][4 * kq + 0][n] = rb.x;
      Bs[buf][4 * kq + 1][n] = rb.y;
      Bs[buf][4 * kq + 2][n] = rb.z;
      Bs[buf][4 * kq + 3][n] = rb.w;
    } else {
      const int k = tid / (BN / 4), nq = tid % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][k][4 * nq]) = rb;
    }
  };

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int n_k = g.K / BK;
  for (int kt = 0; kt < n_k; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_k) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM];
#pragma unroll
      for (int q = 0; q < TM / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4 * q]);
        a[4 * q + 0] = v.x;
        a[4 * q + 1] = v.y;
        a[4 * q + 2] = v.z;
        a[4 * q + 3] = v.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < n_k) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: thread owns rows m0 + ty*TM + i, columns n0 + tx*4 .. +3
  const int n = n0 + tx * 4;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI == EPI_BIAS_ACT) bias4 = *reinterpret_cast<const float4*>(g.bias + n);
  float* Cz = g.C + (EPI == EPI_SPLITK ? blockIdx.z * g.c_slice : 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    float4* dst = reinterpret_cast<float4*>(Cz + static_cast<size_t>(m) * g.ldc + n);
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (EPI == EPI_BIAS_ACT) {
      v.x += bias4.x;
      v.y += bias4.y;
      v.z += bias4.z;
      v.w += bias4.w;
      dispatch_act(g.act, [&](auto tag) {
        constexpr int ACT = decltype(tag)::value;
        v.x = act_fwd_t<ACT>(v.x);
        v.y = act_fwd_t<ACT>(v.y);
        v.z = act_fwd_t<ACT>(v.z);
        v.w = act_fwd_t<ACT>(v.w);
      });
    } else if (EPI == EPI_MUL_ACTGRAD) {
      const float4 h = *reinterpret_cast<const float4*>(g.aux + static_cast<size_t>(m) * g.ldc + n);
      dispatch_act(g.act, [&](auto tag) {
        constexpr int ACT = decltype(tag)::value;
        v.x *= act_grad_out_t<ACT>(h.x);
        v.y *= act_grad_out_t<ACT>(h.y);
        v.z *= act_grad_out_t<ACT>(h.z);
        v.w *= act_grad_out_t<ACT>(h.w);
      });
    } else if (EPI == EPI_ACCUM) {
      const float4 c0 = *dst;
      v.x += c0.x;
      v.y += c0.y;
      v.z += c0.z;
      v.w += c0.w;
    }
    *dst = v;
  }
}

template <bool A_T, bool B_T, int EPI>
static int sgemm(const GemmArgs& g, int splits, void* stream) {
  // BM = 128 when that still yields >= ~1 wave of CTAs, else 64 (small batches / weight-gradient tiles)
  const bool big = (g.M % 128 == 0) && (static_cast<long long>(g.M / 128) * (g.N / BN) * splits >= 148);
  if (big) {
    auto k = sgemm_kernel<128, A_T, B_T, EPI>;
    HDPO_LAUNCH_PDL(k, dim3(g.N / BN, g.M / 128, splits), GEMM_THREADS, 0, stream, g);
  } else {
    auto k = sgemm_kernel<64, A_T, B_T, EPI>;
    HDPO_LAUNCH_PDL(k, dim3(g.N / BN, g.M / 64, splits), GEMM_THREADS, 0, stream, g);
  }
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

// ------------------------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------------------------

// flat state_dict parameters -> zero-padded [Np][Kp] weight slab + [Np] bias
__device__ __forceinline__ float tf32_round(float x) {
#ifdef HDPO_EMU
  return x;
#else
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#endif
}

// tensor-core mode: W_lo, WT (transposed [Kp][Np]) and WT_lo are non-null and Wp receives the hi half
__global__ void __launch_bounds__(256) pack_layer_kernel(const float* __restrict__ params, int gw, int gb, int N, int K,
                                                         int Np, int Kp, float* __restrict__ Wp, float* __restrict__ bp,
                                                         float* __restrict__ W_lo, float* __restrict__ WT,
                                                         float* __restrict__ WT_lo) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np * Kp) {
    const int n = i / Kp, k = i % Kp;
    const float w = (n < N && k < K) ? params[gw + n * K + k] : 0.f;
    if (W_lo) {
      const float hi = tf32_round(w), lo = tf32_round(w - hi);
      Wp[i] = hi;
      W_lo[i] = lo;
      WT[static_cast<size_t>(k) * Np + n] = hi;
      WT_lo[static_cast<size_t>(k) * Np + n] = lo;
    } else {
      Wp[i] = w;
    }
  }
  if (i < Np) bp[i] = i < N ? params[gb + i] : 0.f;
}

// row-wise split of a [rows][ld] fp32 array into (hi, lo)
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, size_t n) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i], h = tf32_round(v);
  hi[i] = h;
  lo[i] = tf32_round(v - h);
}

// initial state rows X_0[b] = [store inventories flat | warehouse inventories flat | 0 pad]; padded rows = 0
__global__ void __launch_bounds__(256) init_state_kernel(const float* __restrict__ store, const float* __restrict__ wh,
                                                         int B, int Bp, int nS, int nW, int ldx, float* __restrict__ X,
                                                         float* __restrict__ cost_b, float* __restrict__ report_b) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<size_t>(Bp) * ldx) {
    const int b = static_cast<int>(i / ldx), k = static_cast<int>(i % ldx);
    float v = 0.f;
    if (b < B) {
      if (k < nS) v = store[static_cast<size_t>(b) * nS + k];
      else if (k < nS + nW) v = wh[static_cast<size_t>(b) * nW + (k - nS)];
    }
    X[i] = v;
  }
  if (i < static_cast<size_t>(B)) {
    cost_b[i] = 0.f;
    if (report_b) report_b[i] = 0.f;
  }
}

__global__ void __launch_bounds__(256) zero_kernel(float* __restrict__ p, size_t n) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// final state rows -> reference layouts
__global__ void __launch_bounds__(256) export_state_kernel(const float* __restrict__ X, int B, int nS, int nW, int ldx,
                                                           float* __restrict__ store, float* __restrict__ wh) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * (nS + nW)) return;
  const int b = static_cast<int>(i / (nS + nW)), k = static_cast<int>(i % (nS + nW));
  const float v = X[static_cast<size_t>(b) * ldx + k];
  if (k < nS) {
    if (store) store[static_cast<size_t>(b) * nS + k] = v;
  } else if (wh) {
    wh[static_cast<size_t>(b) * nW + (k - nS)] = v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// warehouse policy head + simulator period: one WARP per scenario, lanes over stores
// ------------------------------------------------------------------------------------------------------------
struct HeadArgs {
  int B, Bp, S, W, L, Lw, T_stride, tt;  // tt = t + period_shift (demand column); Bp = rows incl. tile padding
  int ldx, ldy, demand_layout;
  int demand_bstride;  // HDPO_DEMAND_TSB: scenarios in the WHOLE batch (this launch may cover a chunk of them)
  int lost, profit, has_edge, transshipment, discrete, in_report;
  float wub;
  const int32_t* adjacency;  // [W][S] or null
  const float* demands;
  HdpoStatics st;
  TraceRef trace;
  // split-K partials of the thin chain GEMMs (n_part <= 1: none). Forward head: Y row = sum of y_nz partial rows
  // (stride part_stride floats), written to y_out; adjoint head: entry gX row += sum of gx_nz partial rows.
  const float* part;
  size_t part_stride;
  int n_part;
  float* y_out;
};


// HDPO_HEAD_PDL: which head kernels let the dependent kernel of their stream (the next tile GEMM) start its prologue as
// soon as they are past their own dependency wait (a GEMM CTA fits next to the small head CTAs of an SM): bit 0 = forward
// head, bit 1 = adjoint head. Measured on B200 (8192 x 50 x 50 stores, three interleaved runs): both heads: forward 5.38 ->
// 5.30 ms, adjoint 10.68 -> 10.90 ms; the default is the forward head only.
#ifndef HDPO_HEAD_PDL
#define HDPO_HEAD_PDL 1
#endif
template <int BIT>
__device__ __forceinline__ void head_pdl_trigger() {
#if !defined(HDPO_EMU)
  if ((HDPO_HEAD_PDL >> BIT) & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ float demand_of(const HeadArgs& a, int b, int s) {
  if (a.demand_layout == HDPO_DEMAND_TSB) return __ldg(a.demands + (static_cast<size_t>(a.tt) * a.S + s) * a.demand_bstride + b);
  return __ldg(a.demands + (static_cast<size_t>(b) * a.S + s) * a.T_stride + a.tt);
}

// softmax shares p[s*W+w] of warehouse w over its connected stores (+ constant hold logit 1.0 unless transshipment)
// written to `share` (shared memory scratch of this warp). neural_networks.py:140-166, 399-419.
__device__ __forceinline__ void softmax_shares(const HeadArgs& a, const float* __restrict__ y, float* __restrict__ share,
                                               int lane) {
  for (int w = 0; w < a.W; ++w) {
    float mx = a.transshipment ? -INFINITY : 1.f;
    for (int s = lane; s < a.S; s += 32) {
      const bool conn = !a.adjacency || a.adjacency[w * a.S + s] != 0;
      if (conn) mx = fmaxf(mx, y[s * a.W + w]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < a.S; s += 32) {
      const bool conn = !a.adjacency || a.adjacency[w * a.S + s] != 0;
      const float e = conn ? expf(y[s * a.W + w] - mx) : 0.f;
      share[s * a.W + w] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (!a.transshipment) sum += expf(1.f - mx);
    const float inv = 1.f / sum;
    for (int s = lane; s < a.S; s += 32) share[s * a.W + w] *= inv;
  }
  __syncwarp();
}

// Per-warp shared-memory rows. Everything a scenario needs is fetched with ONE round of independent, coalesced loads
// (state row, policy output row, lead times, cost coefficients, demands), the period is computed in shared memory,
// and the result rows go back with coalesced float4 stores: the kernels sit on the dependent chain of every period,
// so their latency (not their throughput) is what matters.
struct HeadSmem {
  float *share, *xs, *xo, *y, *lt, *h, *p, *d;
};
__device__ __forceinline__ HeadSmem head_smem_rows(float* base, const HeadArgs& a, bool bwd) {
  const int SW = a.S * a.W;
  HeadSmem r;
  // rows that are accessed as float4 first (the dynamic shared memory base is 16-byte aligned, ldx / ldy % 4 == 0)
  r.xs = base;
  r.xo = r.xs + a.ldx;
  r.y = r.xo + a.ldx;
  float* q = r.y + (bwd ? 2 : 1) * a.ldy;  // bwd: y row followed by the gy row
  r.share = q;
  q += bwd ? 2 * SW + 64 : SW + 32;
  r.lt = q;
  q += SW;
  r.h = q;
  q += a.S;
  r.p = q;
  q += a.S;
  r.d = q;
  return r;
}
// stage the rows of scenario b (all loads are independent of each other). `x` (and `y` when given) must not be
// produced by the kernel's immediate predecessor when this is called before griddepcontrol.wait.
__device__ __forceinline__ void head_stage(const HeadArgs& a, const HeadSmem& r, const float* __restrict__ x,
                                           const float* __restrict__ y, int b, int lane) {
  for (int k = lane * 4; k < a.ldx; k += 128)
    *reinterpret_cast<float4*>(r.xs + k) = *reinterpret_cast<const float4*>(x + k);
  if (y)
    for (int k = lane * 4; k < a.ldy; k += 128)
      *reinterpret_cast<float4*>(r.y + k) = *reinterpret_cast<const float4*>(y + k);
  const int SW = a.S * a.W;
  const float* lt = a.st.lead_times + static_cast<size_t>(b) * SW;
  for (int k = lane; k < SW; k += 32) r.lt[k] = __ldg(lt + k);
  const float* hc = a.st.holding_costs + static_cast<size_t>(b) * a.S;
  const float* pc = a.st.underage_costs + static_cast<size_t>(b) * a.S;
  for (int s = lane; s < a.S; s += 32) {
    r.h[s] = __ldg(hc + s);
    r.p[s] = __ldg(pc + s);
    r.d[s] = demand_of(a, b, s);
  }
}

__global__ void __launch_bounds__(HEAD_WARPS * 32)
warehouse_head_fwd_kernel(HeadArgs a, const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ Xn,
                          float* __restrict__ cost_b, float* __restrict__ report_b, float* __restrict__ reward_t,
                          float* __restrict__ Xn_hi, float* __restrict__ Xn_lo) {
  // Launched with PDL: everything up to pdl_wait() overlaps the predecessor (the output-layer GEMM that writes Y).
  // X_t, the cost coefficients, lead times and demands were produced at least two kernels back, so their (cold,
  // HBM-latency) loads are issued first and only the Y row waits for the predecessor.
  HDPO_DYN_SMEM(float, smem);
  struct TraceScope {  // one record per CTA, emitted when the first thread leaves the kernel body
    const TraceRef& tr;
    unsigned long long t0;
    unsigned int staged_ns;  // time until the staged rows were readable (bits 32.. of the block field)
    __device__ ~TraceScope() {
      if (threadIdx.x == 0) trace_emit(tr, t0, blockIdx.x, staged_ns);
    }
  } trace_scope{a.trace, (a.trace.buf && threadIdx.x == 0) ? trace_now() : 0ull, 0u};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * HEAD_WARPS + warp;
  if (b >= a.Bp) return;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b >= a.B) {  // tile-padding rows stay exactly zero so that they never contribute to a weight gradient
    pdl_wait();
    for (int k = lane * 4; k < a.ldx; k += 128) {
      *reinterpret_cast<float4*>(Xn + static_cast<size_t>(b) * a.ldx + k) = zero4;
      if (Xn_hi) {
        *reinterpret_cast<float4*>(Xn_hi + static_cast<size_t>(b) * a.ldx + k) = zero4;
        *reinterpret_cast<float4*>(Xn_lo + static_cast<size_t>(b) * a.ldx + k) = zero4;
      }
    }
    return;
  }
  const int SW = a.S * a.W;
  const int nS = a.S * a.L;
  const HeadSmem r = head_smem_rows(smem + warp * head_smem_floats(a.S, a.W, a.ldx, a.ldy, false), a, false);
  float* share = r.share;  // alloc[s*W+w]
  head_stage(a, r, X + static_cast<size_t>(b) * a.ldx, nullptr, b, lane);
  // statics of the warehouses (lane w)
  float wh_hold = 0.f, wh_edge = 0.f, wh_lead = 0.f;
  if (lane < a.W) {
    const int bw = b * a.W + lane;
    wh_hold = __ldg(a.st.warehouse_holding_costs + bw);
    wh_lead = __ldg(a.st.warehouse_lead_times + bw);
    if (a.has_edge) wh_edge = __ldg(a.st.warehouse_edge_costs + bw);
  }
  pdl_wait();
  head_pdl_trigger<0>();
  trace_scope.t0 = (a.trace.buf && threadIdx.x == 0) ? trace_now() : 0ull;  // records start after the dependency wait
  if (a.n_part > 1) {
    // split-K output layer: the policy outputs are the sum of the partial rows (fixed order); the row also goes to the
    // Y tape, which the adjoint head reads
    for (int k = lane * 4; k < a.ldy; k += 128) {
      float4 v = *reinterpret_cast<const float4*>(a.part + static_cast<size_t>(b) * a.ldy + k);
      for (int z = 1; z < a.n_part; ++z) {
        const float4 u = *reinterpret_cast<const float4*>(a.part + z * a.part_stride + static_cast<size_t>(b) * a.ldy + k);
        v.x += u.x;
        v.y += u.y;
        v.z += u.z;
        v.w += u.w;
      }
      *reinterpret_cast<float4*>(r.y + k) = v;
      *reinterpret_cast<float4*>(a.y_out + static_cast<size_t>(b) * a.ldy + k) = v;
    }
  } else {
    const float* yrow = Y + static_cast<size_t>(b) * a.ldy;
    for (int k = lane * 4; k < a.ldy; k += 128)
      *reinterpret_cast<float4*>(r.y + k) = *reinterpret_cast<const float4*>(yrow + k);
  }
  __syncwarp();
  if (a.trace.buf && threadIdx.x == 0) {
    volatile float probe = r.xs[0] + r.d[0] + r.h[0];  // force the staged loads to have landed
    (void)probe;
    trace_scope.staged_ns = static_cast<unsigned int>(trace_now() - trace_scope.t0);
  }
  const float* x = r.xs;
  const float* y = r.y;
  float* xn = r.xo;
  softmax_shares(a, y, share, lane);
  // allocations = share * warehouse on-hand (optionally rounded half-to-even)
  for (int i = lane; i < SW; i += 32) {
    const int w = i % a.W;
    float al = share[i] * x[nS + w * a.Lw];
    if (a.discrete) al = rintf(al);
    share[i] = al;
  }
  __syncwarp();
  // ---- stores
  float cost = 0.f;
  for (int s = lane; s < a.S; s += 32) {
    const float* xs = x + s * a.L;
    float* xo = xn + s * a.L;
    const float on_hand = xs[0];
    const float d = r.d[s];
    const float raw = on_hand - d;
    const float h = r.h[s], p = r.p[s];
    cost += a.profit ? (-p * fminf(on_hand, d) + h * relu0(raw)) : (p * relu0(-raw) + h * relu0(raw));
    const float post = a.lost ? relu0(raw) : raw;
    xo[0] = post + xs[1];
    for (int k = 1; k < a.L - 1; ++k) xo[k] = xs[k + 1];
    xo[a.L - 1] = 0.f;
    for (int w = 0; w < a.W; ++w) {
      const float al = share[s * a.W + w];
      if (al != 0.f) {
        const int slot = static_cast<int>(r.lt[s * a.W + w]) - 1;
        if (slot >= 0 && slot < a.L) xo[slot] += al;
      }
    }
  }
  // ---- warehouses: lane w (the draw-down of every warehouse is reduced by the whole warp first)
  float drawn = 0.f;
  for (int w = 0; w < a.W; ++w) {
    float part = 0.f;
    for (int s = lane; s < a.S; s += 32) part += share[s * a.W + w];
    part = warp_sum(part);
    if (lane == w) drawn = part;
  }
  if (lane < a.W) {
    const int w = lane;
    const float* xw = x + nS + w * a.Lw;
    float* xo = xn + nS + w * a.Lw;
    const float raw = xw[0] - drawn;
    float aw = sigmoid_f(y[SW + w]) * a.wub;
    if (a.discrete) aw = rintf(aw);
    float cw = wh_hold * relu0(raw);
    if (a.has_edge) cw += wh_edge * aw;
    cost += cw;
    xo[0] = raw + xw[1];
    for (int k = 1; k < a.Lw - 1; ++k) xo[k] = xw[k + 1];
    xo[a.Lw - 1] = 0.f;
    if (aw != 0.f) {
      const int slot = static_cast<int>(wh_lead) - 1;
      if (slot >= 0 && slot < a.Lw) xo[slot] += aw;
    }
  }
  // padding columns of the next state row stay zero
  for (int k = nS + a.W * a.Lw + lane; k < a.ldx; k += 32) xn[k] = 0.f;
  __syncwarp();
  // next state row (and, tensor-core mode, its (hi, lo) split for the next period's layer-0 GEMM): coalesced stores
  for (int k = lane * 4; k < a.ldx; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xn + k);
    *reinterpret_cast<float4*>(Xn + static_cast<size_t>(b) * a.ldx + k) = v;
    if (Xn_hi) {
      float4 hi, lo;
      hi.x = tf32_round(v.x);
      hi.y = tf32_round(v.y);
      hi.z = tf32_round(v.z);
      hi.w = tf32_round(v.w);
      lo.x = tf32_round(v.x - hi.x);
      lo.y = tf32_round(v.y - hi.y);
      lo.z = tf32_round(v.z - hi.z);
      lo.w = tf32_round(v.w - hi.w);
      *reinterpret_cast<float4*>(Xn_hi + static_cast<size_t>(b) * a.ldx + k) = hi;
      *reinterpret_cast<float4*>(Xn_lo + static_cast<size_t>(b) * a.ldx + k) = lo;
    }
  }
  cost = warp_sum(cost);
  if (lane == 0) {
    cost_b[b] += cost;
    if (report_b && a.in_report) report_b[b] += cost;
    if (reward_t) reward_t[b] = cost;
  }
}

// Adjoint of the head + period. gX: on entry adjoint wrt X_{t+1} row, on exit the direct part of the adjoint wrt X_t.
__global__ void __launch_bounds__(HEAD_WARPS * 32)
warehouse_head_bwd_kernel(HeadArgs a, const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ gX,
                          float* __restrict__ gY, float rb, float* __restrict__ gY_lo) {
  // PDL: X_t, Y_t (forward tapes) and the statics do not depend on the predecessor (the dgrad GEMM that finishes
  // gX of period t+1); they are staged before pdl_wait(), only the gX row is read after it.
  HDPO_DYN_SMEM(float, smem);
  struct TraceScope {
    const TraceRef& tr;
    unsigned long long t0;
    __device__ ~TraceScope() {
      if (threadIdx.x == 0) trace_emit(tr, t0, blockIdx.x);
    }
  } trace_scope{a.trace, (a.trace.buf && threadIdx.x == 0) ? trace_now() : 0ull};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * HEAD_WARPS + warp;
  if (b >= a.Bp) return;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b >= a.B) {
    pdl_wait();
    for (int k = lane * 4; k < a.ldy; k += 128) {
      *reinterpret_cast<float4*>(gY + static_cast<size_t>(b) * a.ldy + k) = zero4;
      if (gY_lo) *reinterpret_cast<float4*>(gY_lo + static_cast<size_t>(b) * a.ldy + k) = zero4;
    }
    return;
  }
  const int SW = a.S * a.W;
  const HeadSmem r = head_smem_rows(smem + warp * head_smem_floats(a.S, a.W, a.ldx, a.ldy, true), a, true);
  float* share = r.share;      // p[s*W+w]
  float* galloc = share + SW;  // adjoint of the allocations
  float* wscr = galloc + SW;   // [0..W): g_raw_w, [W..2W): on-hand W0
  float* g = r.xo;             // adjoint row: staged, updated in place, written back
  float* gy = r.y + a.ldy;
  head_stage(a, r, X + static_cast<size_t>(b) * a.ldx, Y + static_cast<size_t>(b) * a.ldy, b, lane);
  float wh_hold = 0.f, wh_edge = 0.f, wh_lead = 0.f;
  if (lane < a.W) {
    const int bw = b * a.W + lane;
    wh_hold = __ldg(a.st.warehouse_holding_costs + bw);
    wh_lead = __ldg(a.st.warehouse_lead_times + bw);
    if (a.has_edge) wh_edge = __ldg(a.st.warehouse_edge_costs + bw);
  }
  pdl_wait();
  head_pdl_trigger<1>();
  trace_scope.t0 = (a.trace.buf && threadIdx.x == 0) ? trace_now() : 0ull;
  float* gx_row = gX + static_cast<size_t>(b) * a.ldx;
  for (int k = lane * 4; k < a.ldx; k += 128) {
    float4 v = *reinterpret_cast<const float4*>(gx_row + k);
    // split-K first-layer dgrad of the period before (t + 1): its partial products complete the adjoint wrt X_{t+1}
    for (int z = 0; z < a.n_part; ++z) {
      const float4 u = *reinterpret_cast<const float4*>(a.part + z * a.part_stride + static_cast<size_t>(b) * a.ldx + k);
      v.x += u.x;
      v.y += u.y;
      v.z += u.z;
      v.w += u.w;
    }
    *reinterpret_cast<float4*>(g + k) = v;
  }
  __syncwarp();
  const float* x = r.xs;
  const float* y = r.y;
  const int nS = a.S * a.L;
  softmax_shares(a, y, share, lane);
  // ---- warehouses first (lane w): g_raw_w feeds the store-allocation adjoints
  float drawn = 0.f;
  for (int w = 0; w < a.W; ++w) {
    const float W0w = x[nS + w * a.Lw];
    float part = 0.f;
    for (int s = lane; s < a.S; s += 32) part += share[s * a.W + w] * W0w;
    part = warp_sum(part);
    if (lane == w) drawn = part;
  }
  if (lane < a.W) {
    const int w = lane;
    const float W0 = x[nS + w * a.Lw];
    const float raw = W0 - drawn;
    float* gw = g + nS + w * a.Lw;
    const float sg = sigmoid_f(y[SW + w]);
    const float aw = sg * a.wub;
    float gaw = 0.f;
    if (aw != 0.f) {
      const int slot = static_cast<int>(wh_lead) - 1;
      if (slot >= 0 && slot < a.Lw) gaw = gw[slot];
    }
    if (a.has_edge) gaw += rb * wh_edge;
    const float gn0 = gw[0];
    const float g_raw = rb * wh_hold * ge0(raw) + gn0;
    for (int k = a.Lw - 1; k >= 2; --k) gw[k] = gw[k - 1];
    gw[1] = gn0;
    gw[0] = g_raw;
    wscr[w] = g_raw;
    wscr[a.W + w] = W0;
    gy[SW + w] = gaw * a.wub * sg * (1.f - sg);
  }
  __syncwarp();
  // ---- stores (lane s): dynamics adjoint + allocation adjoints
  for (int s = lane; s < a.S; s += 32) {
    const float* xs = x + s * a.L;
    float* gs = g + s * a.L;
    const float on_hand = xs[0];
    const float d = r.d[s];
    const float raw = on_hand - d;
    const float h = r.h[s], p = r.p[s];
    for (int w = 0; w < a.W; ++w) {
      const float al = share[s * a.W + w] * wscr[a.W + w];
      float ga = 0.f;
      if (al != 0.f) {
        const int slot = static_cast<int>(r.lt[s * a.W + w]) - 1;
        if (slot >= 0 && slot < a.L) ga = gs[slot];
      }
      galloc[s * a.W + w] = ga - wscr[w];
    }
    float g0;
    if (a.profit) {
      const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
      g0 = rb * (-p * tie + h * ge0(raw));
    } else {
      g0 = rb * (-p * le0(raw) + h * ge0(raw));
    }
    const float gn0 = gs[0];
    g0 += a.lost ? gn0 * ge0(raw) : gn0;
    for (int k = a.L - 1; k >= 2; --k) gs[k] = gs[k - 1];
    gs[1] = gn0;
    gs[0] = g0;
  }
  __syncwarp();
  // ---- policy head adjoint: alloc = p * W0 ; softmax backward over the connected set
  for (int w = 0; w < a.W; ++w) {
    const float W0 = wscr[a.W + w];
    float dot = 0.f;  // sum_s g_alloc * p   (= adjoint of W0, and with W0 the softmax inner product)
    for (int s = lane; s < a.S; s += 32) dot += galloc[s * a.W + w] * share[s * a.W + w];
    dot = warp_sum(dot);
    if (lane == 0) g[nS + w * a.Lw] += dot;
    for (int s = lane; s < a.S; s += 32) {
      const float pw = share[s * a.W + w];
      gy[s * a.W + w] = pw * (galloc[s * a.W + w] * W0 - dot * W0);
    }
  }
  for (int k = SW + a.W + lane; k < a.ldy; k += 32) gy[k] = 0.f;
  __syncwarp();
  // write back: adjoint row of X_t (direct part) and gY (tensor-core mode: hi half in place, remainder in gY_lo)
  for (int k = lane * 4; k < a.ldx; k += 128)
    *reinterpret_cast<float4*>(gx_row + k) = *reinterpret_cast<const float4*>(g + k);
  for (int k = lane * 4; k < a.ldy; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(gy + k);
    if (gY_lo) {
      float4 hi, lo;
      hi.x = tf32_round(v.x);
      hi.y = tf32_round(v.y);
      hi.z = tf32_round(v.z);
      hi.w = tf32_round(v.w);
      lo.x = tf32_round(v.x - hi.x);
      lo.y = tf32_round(v.y - hi.y);
      lo.z = tf32_round(v.z - hi.z);
      lo.w = tf32_round(v.w - hi.w);
      *reinterpret_cast<float4*>(gY + static_cast<size_t>(b) * a.ldy + k) = hi;
      *reinterpret_cast<float4*>(gY_lo + static_cast<size_t>(b) * a.ldy + k) = lo;
    } else {
      *reinterpret_cast<float4*>(gY + static_cast<size_t>(b) * a.ldy + k) = v;
    }
  }
}

// column sums of a [rows][ld] matrix (bias gradients), two deterministic stages. Stage 1: CTA (x, chunk) sums 64
// columns (16 float4 groups) over the rows of its chunk with 16 row lanes, then a fixed-order shared-memory reduction.
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const float* __restrict__ G, const float* __restrict__ G2,
                                                            size_t rows, int ld, int n_chunks,
                                                            float* __restrict__ part) {
  pdl_wait();
  __shared__ float4 red[16][16];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int col = blockIdx.x * 64 + tx * 4;
  const int chunk = blockIdx.y;
  const size_t per = (rows + n_chunks - 1) / n_chunks;
  const size_t r0 = chunk * per, r1 = (r0 + per < rows) ? r0 + per : rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < ld) {
    for (size_t r = r0 + ty; r < r1; r += 16) {
      float4 v = *reinterpret_cast<const float4*>(G + r * ld + col);
      if (G2) {
        const float4 w = *reinterpret_cast<const float4*>(G2 + r * ld + col);
        v.x += w.x;
        v.y += w.y;
        v.z += w.z;
        v.w += w.w;
      }
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < ld) {
    for (int i = 1; i < 16; ++i) {
      const float4 v = red[i][tx];
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    *reinterpret_cast<float4*>(part + static_cast<size_t>(chunk) * ld + col) = s;
  }
}

// grad[gw + n*K + k] = sum_z part[z][n][k] ; grad[gb + n] = sum_chunks bpart[chunk][n]
// `transposed`: the partial slices hold dW^T ([Kp rows][Np cols], leading dimension ldp = Np) instead of dW ([..][Kp])
// Rows n0 .. n0+N-1 of the computed gradient go to grad[gw + n*dst_ld + k] (dst_ld = K for a plain layer; the
// SymmetryAware projection scatters its two row blocks into the context columns of the store / warehouse nets).
__global__ void __launch_bounds__(256) unpack_grad_kernel(const float* __restrict__ part, int splits, size_t slice, int ldp,
                                                          int transposed, int N, int K, const float* __restrict__ bpart,
                                                          int n_chunks, int ldb, int gw, int gb,
                                                          float* __restrict__ grad, int n0, int dst_ld) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * K) {
    const int n = i / K, k = i % K;
    const size_t at = transposed ? static_cast<size_t>(k) * ldp + (n0 + n) : static_cast<size_t>(n0 + n) * ldp + k;
    double s = 0.0;  // hundreds of partial slices: accumulate in double so the reduction adds no fp32 error
    for (int z = 0; z < splits; ++z) s += static_cast<double>(part[z * slice + at]);
    grad[gw + n * dst_ld + k] = static_cast<float>(s);
  }
  if (i < N) {
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += bpart[static_cast<size_t>(c) * ldb + n0 + i];
    grad[gb + i] = s;
  }
}

__global__ void __launch_bounds__(1024) totals_kernel(const float* __restrict__ cost_b, const float* __restrict__ report_b,
                                                      int B, double* __restrict__ totals) {
  pdl_wait();
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
// host orchestration
// ------------------------------------------------------------------------------------------------------------
// The batch is cut into independent scenario CHUNKS that run on concurrent streams. Every kernel of the period chain
// is short (15-45 us) and ends with a drained tail (partial second wave, un-overlapped epilogue, warp-per-scenario
// heads), and consecutive kernels of one chain are strictly dependent; two or more independent chains keep every SM
// fed. Chunks share nothing but the (read-only) parameters; the parameter gradient is their fixed-order sum.
constexpr int kMaxChunks = 8;
constexpr int kChunkMinRows = 2048;

struct Chunking {
  int n, P;
  int b0[kMaxChunks], rows[kMaxChunks];
  size_t ws_off[kMaxChunks];  // byte offset of each chunk's private workspace
  size_t grad_off;            // (n - 1) * P floats: gradients of chunks 1.. before the fixed-order sum
  size_t total;
};

static int requested_chunks(int B, bool sym) {
#ifdef HDPO_EMU
  (void)B;
  (void)sym;
  return 1;  // the host-thread emulator has no streams
#else
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("HDPO_WIDE_CHUNKS");
    env = e ? atoi(e) : -1;
  }
  // SymmetryAware: the per-scenario head kernels dominate and are persistent over the whole machine, so concurrent
  // chunks only make them compete for the SMs (measured 27.4 ms with one chunk vs 28.4 ms with four at B = 8192)
  int n = env > 0 ? env : (sym ? 1 : (B / kChunkMinRows > 1 ? B / kChunkMinRows : 1));
  if (n > kMaxChunks) n = kMaxChunks;
  while (n > 1 && B / n < kRowPad) --n;
  return n;
#endif
}

static Chunking make_chunking(const HdpoRolloutDesc* d) {
  Chunking c;
  const int B = d->pb.B;
  const int want = use_persist(d) ? 1 : requested_chunks(B, d->arch == HDPO_ARCH_SYMMETRY_AWARE);
  const int per = pad_to((B + want - 1) / want, kRowPad);
  c.n = 0;
  size_t o = 0;
  for (int i = 0; i < want; ++i) {
    const int b0 = i * per;
    const int rows = (B - b0 < per) ? B - b0 : per;
    if (rows <= 0 && i > 0) break;
    c.b0[i] = b0;
    c.rows[i] = rows > 0 ? rows : 0;
    c.ws_off[i] = o;
    const Plan p = make_plan(d, c.rows[i]);
    o += a256(p.total);
    c.P = p.P;
    c.n = i + 1;
  }
  c.grad_off = o;
  o += a256(static_cast<size_t>(c.n > 1 ? c.n - 1 : 0) * c.P * sizeof(float));
  c.total = o + 256;
  return c;
}

size_t workspace_bytes(const HdpoRolloutDesc* d) { return make_chunking(d).total; }

#ifndef HDPO_EMU
// side streams of the current device (created once); serialised by a mutex so that two host threads forking at the
// same time cannot interleave their event records
struct SideStreams {
  cudaStream_t s[kMaxChunks - 1];
  cudaEvent_t fork, join[kMaxChunks - 1];
  // overlapped weight gradients: one chain stream per chunk at the highest priority (chunk 0 leaves the caller's
  // stream too, so that no chunk is scheduled behind the others) and one low-priority stream per chunk for the
  // weight-gradient groups; ev_rows = "the gz rows of the group are written", ev_done = "all groups finished"
  cudaStream_t hi[kMaxChunks], wg[kMaxChunks];
  cudaEvent_t join_hi[kMaxChunks], ev_rows[kMaxChunks], ev_done[kMaxChunks];
  bool ready;
};
static std::mutex g_side_mutex;
static SideStreams g_side[64];

static int get_side_streams(SideStreams** out) {
  int dev = 0;
  HDPO_CUDA_OK(cudaGetDevice(&dev));
  HDPO_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  SideStreams& ss = g_side[dev];
  if (!ss.ready) {
    int prio_low = 0, prio_high = 0;
    HDPO_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
    // HDPO_WIDE_PRIO (A/B knob): 0 = every chunk stream at the default priority, 1 = chunk i at priority -(i) (chunk 0 is
    // the caller's stream), 2 = every side stream one level above the caller's
    const char* pe = getenv("HDPO_WIDE_PRIO");
    const int prio_mode = pe ? atoi(pe) : 0;
    for (int i = 0; i < kMaxChunks - 1; ++i) {
      int pr = prio_low;
      if (prio_mode == 1) pr = prio_low - (i + 1);
      if (prio_mode == 2) pr = prio_low - 1;
      if (pr < prio_high) pr = prio_high;
      HDPO_CUDA_OK(cudaStreamCreateWithPriority(&ss.s[i], cudaStreamNonBlocking, pr));
      HDPO_CUDA_OK(cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming));
    }
    HDPO_CUDA_OK(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i) {
      // (HDPO_WIDE_PRIO = 3: chains and weight-gradient groups at the SAME priority)
      HDPO_CUDA_OK(cudaStreamCreateWithPriority(&ss.hi[i], cudaStreamNonBlocking, prio_mode == 3 ? prio_low : prio_high));
      HDPO_CUDA_OK(cudaStreamCreateWithPriority(&ss.wg[i], cudaStreamNonBlocking, prio_low));
      HDPO_CUDA_OK(cudaEventCreateWithFlags(&ss.join_hi[i], cudaEventDisableTiming));
      HDPO_CUDA_OK(cudaEventCreateWithFlags(&ss.ev_rows[i], cudaEventDisableTiming));
      HDPO_CUDA_OK(cudaEventCreateWithFlags(&ss.ev_done[i], cudaEventDisableTiming));
    }
    ss.ready = true;
  }
  *out = &ss;
  return HDPO_OK;
}
#endif

static float* wsf(void* ws, size_t off) { return reinterpret_cast<float*>(static_cast<char*>(ws) + off); }
template <typename T>
static T* shifted(T* p, size_t n) {
  return p ? p + n : nullptr;
}

#ifndef HDPO_EMU
// tensor maps of one (hi, lo) operand pair
struct MapPair {
  CUtensorMap hi, lo;
};
static int make_pair(MapPair* m, const float* hi, const float* lo, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  int rc = tc::make_tensor_map(&m->hi, hi, rows, cols, cols, box_rows);
  if (rc) return rc;
  return tc::make_tensor_map(&m->lo, lo, rows, cols, cols, box_rows);
}
#endif

// one chunk of scenarios [b0, b0 + p.B): every pointer below is already shifted to the chunk's first scenario
struct ChunkCtx {
  Plan p;
  void* ws;
  void* stream;
  int index;  // chunk number (trace tags)
  int n_chunks;  // chunks of this batch (concurrent chains competing for the SMs)
  const float* demands;
  HdpoStatics st;
  HdpoState init, fin;
  float *cost_b, *report_b, *reward_tb;  // reward_tb rows are B_total apart
  float* grad;
  const float* params;  // flat parameter vector (the SymmetryAware heads stage the local nets from it)
  sym::Cfg sc;          // SymmetryAware head configuration (p.sym only)
#ifndef HDPO_EMU
  MapPair mA[HDPO_MAX_LAYERS], mB[HDPO_MAX_LAYERS];  // GEMM operands (forward: layer input / W; adjoint: gz / W^T)
  MapPair mAct[HDPO_MAX_LAYERS];                     // layer-output tapes as 32-row output boxes (lo unused for the last)
  MapPair mGz[HDPO_MAX_LAYERS];                      // pre-activation adjoint tapes as output boxes
  CUtensorMap mGx;                                   // state adjoint [Bp][wp0]
  CUtensorMap mYpart, mGxPart;                       // split-K partial outputs [n_split * Bp][wp_out] / [.. * Bp][wp0]
#endif
};

static void bind_chunk(ChunkCtx* c, const HdpoRolloutDesc* d, const Chunking& ck, int i, const float* demands,
                       const HdpoStatics* st, void* ws, void* stream) {
  const HdpoProblem& pb = d->pb;
  const size_t b0 = static_cast<size_t>(ck.b0[i]), S = pb.S, W = pb.W;
  c->p = make_plan(d, ck.rows[i]);
  c->index = i;
  c->n_chunks = ck.n;
  c->ws = static_cast<char*>(ws) + ck.ws_off[i];
  c->stream = stream;
  c->demands = d->demand_layout == HDPO_DEMAND_TSB ? demands + b0 : demands + b0 * S * d->t_stride;
  c->st = *st;
  c->st.holding_costs = shifted(st->holding_costs, b0 * S);
  c->st.underage_costs = shifted(st->underage_costs, b0 * S);
  c->st.lead_times = shifted(st->lead_times, b0 * S * W);
  c->st.warehouse_lead_times = shifted(st->warehouse_lead_times, b0 * W);
  c->st.warehouse_holding_costs = shifted(st->warehouse_holding_costs, b0 * W);
  c->st.warehouse_edge_costs = shifted(st->warehouse_edge_costs, b0 * W);
  c->st.mean = shifted(st->mean, b0 * S);
  c->st.std = shifted(st->std, b0 * S);
  c->init = HdpoState{nullptr, nullptr, nullptr};
  c->fin = HdpoState{nullptr, nullptr, nullptr};
  c->cost_b = c->report_b = c->reward_tb = c->grad = nullptr;
  c->params = nullptr;
  if (c->p.sym) sym::build_cfg(d, c->p.B, c->p.Bp, c->p.wp[0], c->p.wp[c->p.n], &c->sc);
}

static sym::PeriodArgs sym_period_args(const HdpoRolloutDesc* d, const ChunkCtx& c, int t) {
  sym::PeriodArgs a;
  a.tt = t + d->period_shift;
  a.in_report = t >= d->ignore_periods;
  a.demands = c.demands;
  a.st = c.st;
  return a;
}

static HeadArgs head_args(const HdpoRolloutDesc* d, const ChunkCtx& c, int t) {
  const Plan& p = c.p;
  HeadArgs a;
  a.B = p.B;
  a.Bp = p.Bp;
  a.S = d->pb.S;
  a.W = d->pb.W;
  a.L = d->pb.L;
  a.Lw = d->pb.Lw;
  a.T_stride = d->t_stride;
  a.tt = t + d->period_shift;
  a.ldx = p.wp[0];
  a.ldy = p.wp[p.n];
  a.demand_layout = d->demand_layout;
  a.demand_bstride = d->pb.B;
  a.lost = d->pb.lost_demand;
  a.profit = d->pb.maximize_profit;
  a.has_edge = d->pb.has_edge_cost;
  a.transshipment = d->transshipment;
  a.discrete = d->discrete_allocation;
  a.in_report = t >= d->ignore_periods;
  a.wub = d->warehouse_upper_bound;
  a.adjacency = d->pb.W > 1 ? d->adjacency : nullptr;
  a.demands = c.demands;
  a.st = c.st;
  a.trace = trace_ref(0xF00u | static_cast<unsigned>(c.index));  // tag: head kernel of chunk index
  a.part = nullptr;
  a.part_stride = 0;
  a.n_part = 0;
  a.y_out = nullptr;
  return a;
}

#ifndef HDPO_EMU
// tile form of the forward GEMM of layer l / of the dgrad GEMM that consumes gz_l: hidden-layer epilogues exist in the
// CTA-pair form (256 x 128 tiles), the output layer and the accumulation into gX use single-CTA tiles
// HDPO_TC_BN = 64 | 128 forces single-CTA tiles of that width everywhere (A/B comparison on the GPU box)
static int forced_bn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDPO_TC_BN");
    v = e ? atoi(e) : 0;
  }
  return v;
}
// HDPO_TC_OCC2 = 1: the per-period GEMMs run as 64-column tiles with TWO CTAs per SM (256 TMEM columns and a 2-stage
// ring each; CTA pairs where the epilogue exists in that form), so that kernels of different chunk streams share SMs
static int g_occ2 = -1;  // < 0: not read yet (HDPO_TC_OCC2, default 0)
static int occ2_mode() {
  if (g_occ2 < 0) {
    const char* e = getenv("HDPO_TC_OCC2");
    g_occ2 = e ? atoi(e) : 0;
    if (g_occ2 < 0) g_occ2 = 0;
  }
  return g_occ2;
}
static int tile_bn(int rows, int cols, bool pair_ok) {
  const int f = forced_bn();
  if (f == 64 || (f == 128 && cols % 128 == 0)) return f;
  if (f == 0 && occ2_mode() && cols % 64 == 0) {
    if (pair_ok && rows % 256 == 0 && tc::pair_enabled() && occ2_mode() != 2) return tc::kBnPair64;
    return tc::kBn64x2;
  }
  if (f == 0) {
    const int m = wp::pick_bn_multi(rows, cols);
    if (m) return m;
  }
  // few tiles (e.g. 1024 scenarios per GPU): 128 x 64 tiles double the CTAs of a launch that cannot fill the machine
  // anyway (measured on many_warehouses 3 x 50, 1024 scenarios: 7.77 -> 7.32 ms per step; 2048-row chunks lose 25 %)
  if (f == 0 && pair_ok && (rows / 128) * (cols / 128) <= 32 && cols % 128 == 0) return 64;
  return pair_ok ? tc::pick_bn_pair(rows, cols) : tc::pick_bn(cols);
}
static int fwd_bn(const Plan& p, int l) { return tile_bn(p.Bp, p.wp[l + 1], l + 1 < p.n); }
static int dgrad_bn(const Plan& p, int l) { return tile_bn(p.Bp, p.wp[l], l > 0); }

// output-side tensor maps (32-row boxes) of the activation tapes, and for the adjoint of the gz tapes and gX
static int make_tape_maps(ChunkCtx& c) {
  const Plan& p = c.p;
  void* ws = c.ws;
  const uint64_t tslots = p.save ? static_cast<uint64_t>(p.T) : 1;
  for (int l = 0; l < p.n; ++l) {
    const bool hidden = l + 1 < p.n;
    int rc = make_pair(&c.mAct[l], wsf(ws, p.o_act[l]), hidden ? wsf(ws, p.o_act_lo[l]) : wsf(ws, p.o_act[l]),
                       tslots * p.Bp, p.wp[l + 1], tc::kBoxRowsC);
    if (rc) return rc;
    if (p.save) {
      rc = make_pair(&c.mGz[l], wsf(ws, p.o_gz[l]), wsf(ws, p.o_gz_lo[l]), tslots * p.Bp, p.wp[l + 1], tc::kBoxRowsC);
      if (rc) return rc;
    }
  }
  if (p.y_split > 1) {
    int rc = tc::make_tensor_map(&c.mYpart, wsf(ws, p.o_ypart), static_cast<uint64_t>(p.y_split) * p.Bp, p.wp[p.n], p.wp[p.n],
                                 tc::kBoxRowsC);
    if (rc) return rc;
  }
  if (p.save && p.gx_split > 1) {
    int rc = tc::make_tensor_map(&c.mGxPart, wsf(ws, p.o_gxpart), static_cast<uint64_t>(p.gx_split) * p.Bp, p.wp[0],
                                 p.wp[0], tc::kBoxRowsC);
    if (rc) return rc;
  }
  if (p.save) return tc::make_tensor_map(&c.mGx, wsf(ws, p.o_gx), p.Bp, p.wp[0], p.wp[0], tc::kBoxRowsC);
  return HDPO_OK;
}
#endif

// tile selector of layer l's forward GEMM, 0 when there is no tensor-core path in this build
static int fwd_bn_host(const Plan& p, int l) {
#ifndef HDPO_EMU
  return fwd_bn(p, l);
#else
  (void)p;
  (void)l;
  return 0;
#endif
}
static int dgrad_bn_host(const Plan& p, int l) {
#ifndef HDPO_EMU
  return dgrad_bn(p, l);
#else
  (void)p;
  (void)l;
  return 0;
#endif
}

// ---- forward: prologue (pack weights, initial state), one period, epilogue (final state) of one chunk ----
static int fwd_begin(ChunkCtx& c, const HdpoRolloutDesc* d, const float* params) {
  const Plan& p = c.p;
  void* ws = c.ws;
  void* stream = c.stream;
  const HdpoProblem& pb = d->pb;
  const int nS = pb.S * pb.L, nW = pb.W * pb.Lw;
  const size_t tslots = p.save ? static_cast<size_t>(p.T) : 1;
  // pack weights (tensor-core mode: hi/lo halves and their transposes)
  c.params = params;
  for (int l = 0; l < p.n; ++l) {
    const int cnt = p.wp[l + 1] * p.wp[l];
    if (p.sym && l + 1 == p.n) {
      int rc = sym::pack_projection(c.sc, params, p.wp[l], wsf(ws, p.o_W[l]), wsf(ws, p.o_b[l]),
                                    p.tc ? wsf(ws, p.o_W_lo[l]) : static_cast<float*>(nullptr),
                                    p.tc ? wsf(ws, p.o_WT[l]) : static_cast<float*>(nullptr),
                                    p.tc ? wsf(ws, p.o_WT_lo[l]) : static_cast<float*>(nullptr), stream);
      if (rc) return rc;
      continue;
    }
    auto k = pack_layer_kernel;
    HDPO_LAUNCH_PDL(k, ceil_div(cnt, 256), 256, 0, stream, params, p.gw[l], p.gb[l], p.w[l + 1], p.w[l], p.wp[l + 1],
                    p.wp[l], wsf(ws, p.o_W[l]), wsf(ws, p.o_b[l]),
                    p.tc ? wsf(ws, p.o_W_lo[l]) : static_cast<float*>(nullptr),
                    p.tc ? wsf(ws, p.o_WT[l]) : static_cast<float*>(nullptr),
                    p.tc ? wsf(ws, p.o_WT_lo[l]) : static_cast<float*>(nullptr));
    HDPO_LAUNCH_OK();
  }
  {
    auto k = init_state_kernel;
    const size_t cnt = static_cast<size_t>(p.Bp) * p.wp[0];
    HDPO_LAUNCH_PDL(k, static_cast<unsigned>(ceil_div64(cnt, 256)), 256, 0, stream,
                    static_cast<const float*>(c.init.store), static_cast<const float*>(c.init.warehouse), p.B, p.Bp, nS,
                    nW, p.wp[0], wsf(ws, p.o_X), c.cost_b, c.report_b);
    HDPO_LAUNCH_OK();
    if (p.tc) {
      auto ks = split_rows_kernel;
      HDPO_LAUNCH_PDL(ks, static_cast<unsigned>(ceil_div64(cnt, 256)), 256, 0, stream,
                      static_cast<const float*>(wsf(ws, p.o_X)), wsf(ws, p.o_X_hi), wsf(ws, p.o_X_lo), cnt);
      HDPO_LAUNCH_OK();
    }
  }
#ifndef HDPO_EMU
  if (p.tc) {
    int rc_maps = 0;
    for (int l = 0; l < p.n; ++l) {
      const float* a_hi = (l == 0) ? wsf(ws, p.o_X_hi) : wsf(ws, p.o_act[l - 1]);
      const float* a_lo = (l == 0) ? wsf(ws, p.o_X_lo) : wsf(ws, p.o_act_lo[l - 1]);
      int rc = make_pair(&c.mA[l], a_hi, a_lo, tslots * p.Bp, p.wp[l], tc::kBoxRowsA);
      if (rc) return rc;
      rc = make_pair(&c.mB[l], wsf(ws, p.o_W[l]), wsf(ws, p.o_W_lo[l]), p.wp[l + 1], p.wp[l],
                     tc::b_box_rows(fwd_bn(p, l)));
      if (rc) return rc;
    }
    if ((rc_maps = make_tape_maps(c))) return rc_maps;
  }
#endif
  return HDPO_OK;
}

static int fwd_period(ChunkCtx& c, const HdpoRolloutDesc* d, int t) {
  const Plan& p = c.p;
  void* ws = c.ws;
  void* stream = c.stream;
  const HdpoProblem& pb = d->pb;
  const size_t xs = p.save ? static_cast<size_t>(t) : static_cast<size_t>(t & 1);
  const size_t xn = p.save ? static_cast<size_t>(t + 1) : static_cast<size_t>((t + 1) & 1);
  const size_t as = p.save ? static_cast<size_t>(t) : 0;
  const float* X = wsf(ws, p.o_X) + xs * p.x_stride;
  float* Xn = wsf(ws, p.o_X) + xn * p.x_stride;
  const float* in = X;
  for (int l = 0; l < p.n; ++l) {
    float* out = wsf(ws, p.o_act[l]) + as * p.act_stride[l];
    const int act = p.act[l];
    int rc;
    if (!p.tc) {
      GemmArgs g{};
      g.A = in;
      g.B = wsf(ws, p.o_W[l]);
      g.C = out;
      g.M = p.Bp;
      g.N = p.wp[l + 1];
      g.K = p.wp[l];
      g.lda = p.wp[l];
      g.ldb = p.wp[l];
      g.ldc = p.wp[l + 1];
      g.bias = wsf(ws, p.o_b[l]);
      g.act = act;
      rc = sgemm<false, true, EPI_BIAS_ACT>(g, 1, stream);
    } else {
#ifndef HDPO_EMU
      tc::GemmTcArgs g{};
      g.M = p.Bp;
      g.N = p.wp[l + 1];
      g.K = p.wp[l];
      g.n_pass = p.n_pass;
      g.a_row0 = static_cast<int>(as * p.Bp);
      g.b_row0 = 0;
      g.c_row0 = g.a_row0;
      g.trace.tag = static_cast<unsigned>(c.index);
      g.ldc = p.wp[l + 1];
      g.act = act;
      g.bias = wsf(ws, p.o_b[l]);
      const bool hidden = l + 1 < p.n;
      tc::GemmTcMaps tm{c.mA[l].hi, c.mA[l].lo, c.mB[l].hi, c.mB[l].lo, c.mAct[l].hi, c.mAct[l].lo, c.mAct[l].hi,
                        c.mAct[l].lo};
      const int bn_l = fwd_bn(p, l);
      if (!hidden && p.y_split > 1 && (bn_l == 64 || bn_l == 128)) {
        // thin output layer: split-K partial products into the chunk's partial buffer (the head sums them)
        g.k_per_split = g.K / p.y_split;
        g.c_zrows = p.Bp;
        g.c_row0 = 0;
        tm.c0 = tm.c1 = tm.x0 = tm.x1 = c.mYpart;
      }
      rc = tc::gemm(tm, g, hidden ? tc::EPI_FWD_HIDDEN : tc::EPI_FWD_OUT, bn_l, stream);
#else
      rc = HDPO_E_INVALID;
#endif
    }
    if (rc) return rc;
    in = out;
  }
  // tensor-core mode: the state written for period t+1 is also split into the (hi, lo) tape slot t+1
  float* xn_hi = nullptr;
  float* xn_lo = nullptr;
  if (p.tc && (t + 1 < p.T)) {
    const size_t slot = p.save ? static_cast<size_t>(t + 1) : 0;
    xn_hi = wsf(ws, p.o_X_hi) + slot * p.x_stride;
    xn_lo = wsf(ws, p.o_X_lo) + slot * p.x_stride;
  }
  if (p.sym) {
    const sym::PeriodArgs sa = sym_period_args(d, c, t);
    return sym::head_fwd(c.sc, sa, c.params, X, in, Xn, xn_hi, xn_lo,
                         p.save ? wsf(ws, p.o_so) + static_cast<size_t>(t) * p.so_stride : static_cast<float*>(nullptr),
                         c.cost_b, c.report_b,
                         c.reward_tb ? c.reward_tb + static_cast<size_t>(t) * d->pb.B : static_cast<float*>(nullptr),
                         stream);
  }
  HeadArgs a = head_args(d, c, t);
  if (p.tc && p.y_split > 1 && (fwd_bn_host(p, p.n - 1) == 64 || fwd_bn_host(p, p.n - 1) == 128)) {
    a.part = wsf(ws, p.o_ypart);
    a.part_stride = static_cast<size_t>(p.Bp) * p.wp[p.n];
    a.n_part = p.y_split;
    a.y_out = const_cast<float*>(in);  // the Y tape slot of this period
  }
  const size_t head_smem = static_cast<size_t>(HEAD_WARPS) * head_smem_floats(pb.S, pb.W, p.wp[0], p.wp[p.n], false) * sizeof(float);
  auto k = warehouse_head_fwd_kernel;
#ifndef HDPO_EMU
  if (head_smem > 48 * 1024) HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHeadSmemMax)));
#endif
  HDPO_LAUNCH_PDL(k, ceil_div(p.Bp, HEAD_WARPS), HEAD_WARPS * 32, head_smem, stream, a, X, in, Xn, c.cost_b, c.report_b,
                  c.reward_tb ? c.reward_tb + static_cast<size_t>(t) * d->pb.B : static_cast<float*>(nullptr), xn_hi,
                  xn_lo);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

static int fwd_end(ChunkCtx& c, const HdpoRolloutDesc* d) {
  const Plan& p = c.p;
  const HdpoProblem& pb = d->pb;
  const int nS = pb.S * pb.L, nW = pb.W * pb.Lw;
  if (c.fin.store || c.fin.warehouse) {
    const size_t xf = p.save ? static_cast<size_t>(p.T) : static_cast<size_t>(p.T & 1);
    auto k = export_state_kernel;
    const size_t cnt = static_cast<size_t>(p.B) * (nS + nW);
    HDPO_LAUNCH_PDL(k, static_cast<unsigned>(ceil_div64(cnt, 256)), 256, 0, c.stream,
                    static_cast<const float*>(wsf(c.ws, p.o_X) + xf * p.x_stride), p.B, nS, nW, p.wp[0], c.fin.store,
                    c.fin.warehouse);
    HDPO_LAUNCH_OK();
  }
  return HDPO_OK;
}

#ifndef HDPO_EMU
static bool persist_bwd_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDPO_WIDE_PERSIST_BWD");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// hand the chunk's tapes to the persistent sweeps (wide_persist.cu)
static void fill_persist_ctx(wp::Ctx* pc, const ChunkCtx& c, const HdpoRolloutDesc* d) {
  const Plan& p = c.p;
  void* ws = c.ws;
  wp::Ctx& x = *pc;
  x = wp::Ctx{};
  x.B = p.B;
  x.Bp = p.Bp;
  x.T = p.T;
  x.n = p.n;
  for (int i = 0; i <= p.n; ++i) {
    x.w[i] = p.w[i];
    x.wp[i] = p.wp[i];
  }
  x.save = p.save;
  x.n_pass = p.n_pass;
  x.period_shift = d->period_shift;
  x.ignore_periods = d->ignore_periods;
  x.t_stride = d->t_stride;
  x.demand_layout = d->demand_layout;
  x.B_total = d->pb.B;
  x.S = d->pb.S;
  x.W = d->pb.W;
  x.L = d->pb.L;
  x.Lw = d->pb.Lw;
  x.lost = d->pb.lost_demand;
  x.profit = d->pb.maximize_profit;
  x.has_edge = d->pb.has_edge_cost;
  x.transshipment = d->transshipment;
  x.discrete = d->discrete_allocation;
  x.wub = d->warehouse_upper_bound;
  x.adjacency = d->adjacency;
  for (int l = 0; l < p.n; ++l) {
    x.act[l] = p.act[l];
    x.W_hi[l] = wsf(ws, p.o_W[l]);
    x.W_lo[l] = wsf(ws, p.o_W_lo[l]);
    x.bias[l] = wsf(ws, p.o_b[l]);
    x.WT_hi[l] = wsf(ws, p.o_WT[l]);
    x.WT_lo[l] = wsf(ws, p.o_WT_lo[l]);
    x.act_hi[l] = wsf(ws, p.o_act[l]);
    x.act_lo[l] = l + 1 < p.n ? wsf(ws, p.o_act_lo[l]) : nullptr;
    x.gz_hi[l] = p.save ? wsf(ws, p.o_gz[l]) : nullptr;
    x.gz_lo[l] = p.save ? wsf(ws, p.o_gz_lo[l]) : nullptr;
    x.csum[l] = (p.save && l + 1 < p.n) ? wsf(ws, p.o_csum[l]) : nullptr;
  }
  x.gX = p.save ? wsf(ws, p.o_gx) : nullptr;
  x.X = wsf(ws, p.o_X);
  x.X_hi = wsf(ws, p.o_X_hi);
  x.X_lo = wsf(ws, p.o_X_lo);
  x.demands = c.demands;
  x.st = c.st;
  x.cost_b = c.cost_b;
  x.report_b = c.report_b;
  x.reward_tb = c.reward_tb;
  x.extra = static_cast<char*>(ws) + p.o_extra;
  x.stream = c.stream;
}
#endif

// fork / join of the chunk streams around a region of `stream` (no-ops for a single chunk)
struct StreamFork {
#ifndef HDPO_EMU
  SideStreams* ss = nullptr;
  std::unique_lock<std::mutex> lock;
#endif
  int n = 1;
  bool all_side = false;  // every chunk (also chunk 0) on its own high-priority stream + a weight-gradient stream each
  void* main = nullptr;
  int begin(int n_chunks, void* stream, bool overlap = false) {
    n = n_chunks;
    main = stream;
    all_side = overlap;
#ifndef HDPO_EMU
    if (all_side) {
      lock = std::unique_lock<std::mutex>(g_side_mutex);
      int rc = get_side_streams(&ss);
      if (rc) return rc;
      HDPO_CUDA_OK(cudaEventRecord(ss->fork, static_cast<cudaStream_t>(main)));
      for (int i = 0; i < n; ++i) HDPO_CUDA_OK(cudaStreamWaitEvent(ss->hi[i], ss->fork, 0));
      return HDPO_OK;
    }
    if (n > 1) {
      lock = std::unique_lock<std::mutex>(g_side_mutex);
      int rc = get_side_streams(&ss);
      if (rc) return rc;
      HDPO_CUDA_OK(cudaEventRecord(ss->fork, static_cast<cudaStream_t>(main)));
      for (int i = 1; i < n; ++i) HDPO_CUDA_OK(cudaStreamWaitEvent(ss->s[i - 1], ss->fork, 0));
    }
#endif
    return HDPO_OK;
  }
  void* stream_of(int i) const {
#ifndef HDPO_EMU
    if (all_side) return ss->hi[i];
    if (i > 0) return ss->s[i - 1];
#endif
    (void)i;
    return main;
  }
#ifndef HDPO_EMU
  // chunk i's weight-gradient stream may start on rows the chain stream has written so far
  int rows_ready(int i) {
    HDPO_CUDA_OK(cudaEventRecord(ss->ev_rows[i], ss->hi[i]));
    HDPO_CUDA_OK(cudaStreamWaitEvent(ss->wg[i], ss->ev_rows[i], 0));
    return HDPO_OK;
  }
  void* wg_stream_of(int i) const { return ss->wg[i]; }
  // chunk i's chain stream continues after everything queued on its weight-gradient stream
  int wg_join(int i) {
    HDPO_CUDA_OK(cudaEventRecord(ss->ev_done[i], ss->wg[i]));
    HDPO_CUDA_OK(cudaStreamWaitEvent(ss->hi[i], ss->ev_done[i], 0));
    return HDPO_OK;
  }
#endif
  int end() {
#ifndef HDPO_EMU
    if (all_side) {
      for (int i = 0; i < n; ++i) {
        HDPO_CUDA_OK(cudaEventRecord(ss->join_hi[i], ss->hi[i]));
        HDPO_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(main), ss->join_hi[i], 0));
      }
      return HDPO_OK;
    }
    if (n > 1) {
      for (int i = 1; i < n; ++i) {
        HDPO_CUDA_OK(cudaEventRecord(ss->join[i - 1], ss->s[i - 1]));
        HDPO_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(main), ss->join[i - 1], 0));
      }
    }
#endif
    return HDPO_OK;
  }
};

}  // namespace wide
}  // namespace hdpo
#ifndef HDPO_EMU
// Debug: per-role event trace of the persistent wide sweeps (tools/wp_trace.py): buf = 74 * 4 * cap * 2 uint64.
extern "C" int hdpo_debug_set_wp_trace(unsigned long long* buf, int32_t cap_per_role) {
  hdpo::wp::set_trace(buf, cap_per_role);
  return HDPO_OK;
}
// Switch the opt-in persistent one-launch sweeps of the wide path on / off (default: HDPO_WIDE_PERSIST, off).
extern "C" int hdpo_debug_set_wide_persist(int32_t on) {
  hdpo::wp::set_enabled(on);
  return HDPO_OK;
}
// Routing threshold of the multi-tile CTA-pair GEMM: min_tiles > 0 = fewest tiles of a launch that goes there,
// 0 = never, < 0 = back to the default (HDPO_TC_MULTI / HDPO_TC_MULTI_MIN).
extern "C" int hdpo_debug_set_tc_multi(int32_t min_tiles) {
  hdpo::wp::set_multi_min_tiles(min_tiles);
  return HDPO_OK;
}
// Two-CTAs-per-SM tile forms of the per-period GEMMs: 0 = off (default), 1 = 256 x 64 CTA pairs where possible, 2 = 128 x 64
// single-CTA tiles everywhere, < 0 = back to HDPO_TC_OCC2.
extern "C" int hdpo_debug_set_tc_occ2(int32_t mode) {
  hdpo::wide::g_occ2 = mode < 0 ? -1 : mode;
  return HDPO_OK;
}
// Weight-gradient GEMMs overlapped with the adjoint sweep: mode 1 = on, 0 = off, < 0 = by chunk count (default);
// group = periods per group (<= 0: keep). Changes the workspace size: query hdpo_rollout_workspace_bytes afterwards.
extern "C" int hdpo_debug_set_wide_wg_overlap(int32_t mode, int32_t group) {
  hdpo::wide::g_wg_overlap = mode < 0 ? -1 : (mode != 0);
  if (group > 0) hdpo::wide::g_wg_group = group;
  return HDPO_OK;
}
// Split-K of the thin chain GEMMs (output layer, first-layer dgrad): 1 = on, 0 = off, < 0 = default (on for one chunk).
// Changes the workspace size and the summation order of those two layers (results agree to fp32 rounding).
extern "C" int hdpo_debug_set_wide_ksplit(int32_t mode) {
  hdpo::wide::g_ksplit = mode < 0 ? -1 : (mode != 0);
  return HDPO_OK;
}
#else
extern "C" int hdpo_debug_set_wide_ksplit(int32_t) { return HDPO_E_INVALID; }
extern "C" int hdpo_debug_set_tc_occ2(int32_t) { return HDPO_E_INVALID; }
extern "C" int hdpo_debug_set_wide_wg_overlap(int32_t, int32_t) { return HDPO_E_INVALID; }
extern "C" int hdpo_debug_set_tc_multi(int32_t) { return HDPO_E_INVALID; }
extern "C" int hdpo_debug_set_wp_trace(unsigned long long*, int32_t) { return HDPO_E_INVALID; }
extern "C" int hdpo_debug_set_wide_persist(int32_t) { return HDPO_E_INVALID; }
#endif
namespace hdpo {
namespace wide {

int forward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
            const HdpoState* init, float* cost_b, float* report_b, float* reward_tb, double* totals,
            HdpoState* final_state, void* ws, size_t ws_bytes, void* stream) {
  const Chunking ck = make_chunking(d);
  HDPO_REQUIRE(ws != nullptr, "the wide rollout needs its workspace (hdpo_rollout_workspace_bytes)");
  if (ws_bytes < ck.total) {
    set_error("workspace too small: %zu < %zu", ws_bytes, ck.total);
    return HDPO_E_WORKSPACE;
  }
  HDPO_REQUIRE(d->pb.W == 1 || d->adjacency != nullptr, "warehouse_store_adjacency required for n_warehouses > 1");
  const HdpoProblem& pb = d->pb;
  StreamFork fork;
  int rc = fork.begin(ck.n, stream);
  if (rc) return rc;
  ChunkCtx ctx[kMaxChunks];
  for (int i = 0; i < ck.n; ++i) {
    ChunkCtx& c = ctx[i];
    const size_t b0 = static_cast<size_t>(ck.b0[i]);
    bind_chunk(&c, d, ck, i, demands, st, ws, fork.stream_of(i));
    c.init.store = init->store + b0 * pb.S * pb.L;
    c.init.warehouse = init->warehouse + b0 * pb.W * pb.Lw;
    if (final_state) {
      c.fin.store = shifted(final_state->store, b0 * pb.S * pb.L);
      c.fin.warehouse = shifted(final_state->warehouse, b0 * pb.W * pb.Lw);
    }
    c.cost_b = cost_b + b0;
    c.report_b = shifted(report_b, b0);
    c.reward_tb = shifted(reward_tb, b0);
    if ((rc = fwd_begin(c, d, params))) return rc;
  }
#ifndef HDPO_EMU
  if (ctx[0].p.persist) {
    // ONE launch for all periods (wide_persist.cu); same tapes as the per-period chain below
    wp::Ctx pc;
    fill_persist_ctx(&pc, ctx[0], d);
    if ((rc = wp::forward(pc))) return rc;
  } else
#endif
  // period-major issue order: the chunks advance together, so their kernels interleave on the device
  for (int t = 0; t < d->T; ++t)
    for (int i = 0; i < ck.n; ++i)
      if ((rc = fwd_period(ctx[i], d, t))) return rc;
  for (int i = 0; i < ck.n; ++i)
    if ((rc = fwd_end(ctx[i], d))) return rc;
  if ((rc = fork.end())) return rc;
  if (totals) {
    auto k = totals_kernel;
    HDPO_LAUNCH_PDL(k, 1, 1024, 0, stream, static_cast<const float*>(cost_b), static_cast<const float*>(report_b), pb.B,
                    totals);
    HDPO_LAUNCH_OK();
  }
  return HDPO_OK;
}

// ---- adjoint ----
static int bwd_begin(ChunkCtx& c) {
  const Plan& p = c.p;
  void* ws = c.ws;
  float* gX = wsf(ws, p.o_gx);
  {
    auto k = zero_kernel;
    HDPO_LAUNCH_PDL(k, static_cast<unsigned>(ceil_div64(p.x_stride, 256)), 256, 0, c.stream, gX, p.x_stride);
    HDPO_LAUNCH_OK();
    if (p.gx_split > 1) {  // (the first adjoint head of the sweep adds them to its zero gX row)
      const size_t n = static_cast<size_t>(p.gx_split) * p.x_stride;
      HDPO_LAUNCH_PDL(k, static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, c.stream, wsf(ws, p.o_gxpart), n);
      HDPO_LAUNCH_OK();
    }
  }
#ifndef HDPO_EMU
  if (p.tc) {
    for (int l = 0; l < p.n; ++l) {  // dgrad of layer l: A = gz_l [rows][wp[l+1]], B = W_l^T [wp[l]][wp[l+1]]
      int rc = make_pair(&c.mA[l], wsf(ws, p.o_gz[l]), wsf(ws, p.o_gz_lo[l]), static_cast<uint64_t>(p.T) * p.Bp,
                         p.wp[l + 1], tc::kBoxRowsA);
      if (rc) return rc;
      rc = make_pair(&c.mB[l], wsf(ws, p.o_WT[l]), wsf(ws, p.o_WT_lo[l]), p.wp[l], p.wp[l + 1],
                     tc::b_box_rows(dgrad_bn(p, l)));
      if (rc) return rc;
    }
    int rc = make_tape_maps(c);
    if (rc) return rc;
  }
#endif
  return HDPO_OK;
}

static int bwd_period(ChunkCtx& c, const HdpoRolloutDesc* d, int t, float rb) {
  const Plan& p = c.p;
  void* ws = c.ws;
  void* stream = c.stream;
  const HdpoProblem& pb = d->pb;
  float* gX = wsf(ws, p.o_gx);
  const int last = p.n - 1;
  const size_t head_smem = static_cast<size_t>(HEAD_WARPS) * head_smem_floats(pb.S, pb.W, p.wp[0], p.wp[p.n], true) * sizeof(float);
  const float* X = wsf(ws, p.o_X) + static_cast<size_t>(t) * p.x_stride;
  const float* Y = wsf(ws, p.o_act[last]) + static_cast<size_t>(t) * p.act_stride[last];
  float* gY = wsf(ws, p.o_gz[last]) + static_cast<size_t>(t) * p.act_stride[last];
  float* gY_lo = p.tc ? wsf(ws, p.o_gz_lo[last]) + static_cast<size_t>(t) * p.act_stride[last] : nullptr;
  if (p.sym) {
    const sym::PeriodArgs sa = sym_period_args(d, c, t);
    int rc = sym::head_bwd(c.sc, sa, c.params, X, Y, wsf(ws, p.o_so) + static_cast<size_t>(t) * p.so_stride, gX, gY,
                           gY_lo, rb, wsf(ws, p.o_slab), t == p.T - 1, stream);
    if (rc) return rc;
  } else {
    HeadArgs a = head_args(d, c, t);
    const bool gx_parts = p.tc && p.gx_split > 1 && (dgrad_bn_host(p, 0) == 64 || dgrad_bn_host(p, 0) == 128);
    if (gx_parts) {
      a.part = wsf(ws, p.o_gxpart);
      a.part_stride = p.x_stride;
      a.n_part = p.gx_split;
    }
    auto k = warehouse_head_bwd_kernel;
#ifndef HDPO_EMU
    if (head_smem > 48 * 1024) HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHeadSmemMax)));
#endif
    HDPO_LAUNCH_PDL(k, ceil_div(p.Bp, HEAD_WARPS), HEAD_WARPS * 32, head_smem, stream, a, X, Y, gX, gY, rb, gY_lo);
    HDPO_LAUNCH_OK();
  }
  // dgrad chain: gz_{l-1} = (gz_l W_l) * act'(h_{l-1});  finally gX += gz_0 W_0
  for (int l = last; l >= 0; --l) {
    int rc;
    if (!p.tc) {
      GemmArgs g{};
      g.A = wsf(ws, p.o_gz[l]) + static_cast<size_t>(t) * p.act_stride[l];
      g.B = wsf(ws, p.o_W[l]);
      g.M = p.Bp;
      g.N = p.wp[l];
      g.K = p.wp[l + 1];
      g.lda = p.wp[l + 1];
      g.ldb = p.wp[l];
      g.ldc = p.wp[l];
      g.act = l > 0 ? p.act[l - 1] : HDPO_ACT_NONE;
      if (l > 0) {
        g.C = wsf(ws, p.o_gz[l - 1]) + static_cast<size_t>(t) * p.act_stride[l - 1];
        g.aux = wsf(ws, p.o_act[l - 1]) + static_cast<size_t>(t) * p.act_stride[l - 1];
        rc = sgemm<false, false, EPI_MUL_ACTGRAD>(g, 1, stream);
      } else {
        g.C = gX;
        rc = sgemm<false, false, EPI_ACCUM>(g, 1, stream);
      }
    } else {
#ifndef HDPO_EMU
      tc::GemmTcArgs g{};
      g.M = p.Bp;
      g.N = p.wp[l];
      g.K = p.wp[l + 1];
      g.n_pass = p.n_pass;
      g.a_row0 = t * p.Bp;
      g.b_row0 = 0;
      g.trace.tag = static_cast<unsigned>(c.index);
      g.ldc = p.wp[l];
      g.pdl_late = c.n_chunks > 1;
      g.act = l > 0 ? p.act[l - 1] : HDPO_ACT_NONE;
      if (l > 0) {
        g.c_row0 = g.x_row0 = t * p.Bp;
        g.colsum_part = wsf(ws, p.o_csum[l - 1]);
        tc::GemmTcMaps tm{c.mA[l].hi,      c.mA[l].lo,      c.mB[l].hi,       c.mB[l].lo,
                          c.mGz[l - 1].hi, c.mGz[l - 1].lo, c.mAct[l - 1].hi, c.mAct[l - 1].lo};
        rc = tc::gemm(tm, g, tc::EPI_DGRAD_HIDDEN, dgrad_bn(p, l), stream);
      } else {
        g.c_row0 = g.x_row0 = 0;
        const int bn0 = dgrad_bn(p, l);
        if (p.gx_split > 1 && (bn0 == 64 || bn0 == 128)) {
          // thin first-layer dgrad: split-K partial products; the adjoint head of period t - 1 adds them to its gX row
          g.k_per_split = g.K / p.gx_split;
          g.c_zrows = p.Bp;
          tc::GemmTcMaps tm{c.mA[l].hi, c.mA[l].lo, c.mB[l].hi, c.mB[l].lo, c.mGxPart, c.mGxPart, c.mGxPart, c.mGxPart};
          rc = tc::gemm(tm, g, tc::EPI_STORE, bn0, stream);
        } else {
          tc::GemmTcMaps tm{c.mA[l].hi, c.mA[l].lo, c.mB[l].hi, c.mB[l].lo, c.mGx, c.mGx, c.mGx, c.mGx};
          rc = tc::gemm(tm, g, tc::EPI_DGRAD_ACCUM, bn0, stream);
        }
      }
#else
      rc = HDPO_E_INVALID;
#endif
    }
    if (rc) return rc;
  }
  return HDPO_OK;
}

// weight gradients: dW_l[n][k] = sum over all (t, b) rows of gz_l[row][n] * in_l[row][k], split-K over the rows
#ifndef HDPO_EMU
// does layer l take the tcgen05 MN-major form (straight from the tapes)?
static bool wgrad_tc_layer(const Plan& p, int l) { return p.tc && (p.wp[l + 1] % 128 == 0 || p.wp[l] % 128 == 0); }

// tcgen05 weight-gradient GEMM of layer l over the tape rows of periods [t_lo, t_hi]: partial slices
// [t_lo * Bp / kps, (t_hi + 1) * Bp / kps) of o_part_l[l]. When the output width is not a multiple of the 128-row
// MMA tile (e.g. the 64-padded last layer) dW^T = in^T gz is computed instead and transposed while unpacking.
static int wgrad_tc_rows(ChunkCtx& c, int l, int t_lo, int t_hi, void* stream) {
  const Plan& p = c.p;
  void* ws = c.ws;
  const size_t rows = static_cast<size_t>(p.T) * p.Bp;
  const bool transposed = p.wp[l + 1] % 128 != 0;
  const float* gz_hi = wsf(ws, p.o_gz[l]);
  const float* gz_lo = wsf(ws, p.o_gz_lo[l]);
  const float* in_hi = (l == 0) ? wsf(ws, p.o_X_hi) : wsf(ws, p.o_act[l - 1]);
  const float* in_lo = (l == 0) ? wsf(ws, p.o_X_lo) : wsf(ws, p.o_act_lo[l - 1]);
  MapPair mg, mi;
  int rc = tc::make_tensor_map(&mg.hi, gz_hi, rows, p.wp[l + 1], p.wp[l + 1], 32, true);
  if (!rc) rc = tc::make_tensor_map(&mg.lo, gz_lo, rows, p.wp[l + 1], p.wp[l + 1], 32, true);
  if (!rc) rc = tc::make_tensor_map(&mi.hi, in_hi, rows, p.wp[l], p.wp[l], 32, true);
  if (!rc) rc = tc::make_tensor_map(&mi.lo, in_lo, rows, p.wp[l], p.wp[l], 32, true);
  if (rc) return rc;
  tc::GemmTcArgs g{};
  g.M = transposed ? p.wp[l] : p.wp[l + 1];
  g.N = transposed ? p.wp[l + 1] : p.wp[l];
  const size_t row0 = static_cast<size_t>(t_lo) * p.Bp;
  g.K = static_cast<int>(static_cast<size_t>(t_hi - t_lo + 1) * p.Bp);
  g.k_per_split = p.wg_kps;
  g.a_row0 = g.b_row0 = static_cast<int>(row0);
  g.c_row0 = static_cast<int>(row0 / p.wg_kps) * g.M;  // first partial slice of this row range
  g.c_slice = static_cast<size_t>(p.wp[l + 1]) * p.wp[l];
  g.n_pass = p.n_pass;
  g.trace.tag = static_cast<unsigned>(c.index);
  g.ldc = g.N;
  CUtensorMap mpart;  // the partial slices as one [n_slices * M][N] array
  rc = tc::make_tensor_map(&mpart, wsf(ws, p.o_part_l[l]), static_cast<uint64_t>(p.wg_splits) * g.M, g.N, g.N,
                           tc::kBoxRowsC);
  if (rc) return rc;
  tc::GemmTcMaps tm = transposed ? tc::GemmTcMaps{mi.hi, mi.lo, mg.hi, mg.lo, mpart, mpart, mpart, mpart}
                                 : tc::GemmTcMaps{mg.hi, mg.lo, mi.hi, mi.lo, mpart, mpart, mpart, mpart};
  // CTA-pair tiles (256 x 128 / 256 x 64) where the shape allows: each CTA stages its own 128 m-columns and HALF of the
  // n-columns, a quarter fewer operand bytes per flop than the single-CTA 128 x 128 tiles, which ran at the L2 -> SM cap
  // (1 MB per CTA in 12.7 us on 148 SMs). Measured on B200 (8192 x 50 x 50 stores): adjoint 10.73 -> 10.15 ms with the
  // 256 x 128 form. HDPO_WG_PAIR = 0: single-CTA tiles (A/B), 2: pairs for the 128-multiple widths only.
  static int wg_pair = -1;
  if (wg_pair < 0) {
    const char* e = getenv("HDPO_WG_PAIR");
    wg_pair = e ? atoi(e) : 1;
  }
  int bn_w = tc::pick_bn(g.N);
  if (wg_pair && g.M % 256 == 0) {
    if (g.N % 128 == 0) bn_w = tc::kBnPair;
    else if (wg_pair == 1 && g.N % 64 == 0) bn_w = tc::kBnPair64;
  }
  return tc::gemm_wgrad(tm, g, bn_w, stream);
}

// overlapped form: the weight gradients of periods [t_lo, t_hi] of every tcgen05 layer, on the chunk's second stream
static int wgrad_group(ChunkCtx& c, int t_lo, int t_hi, void* wg_stream) {
  for (int l = 0; l < c.p.n; ++l) {
    if (!wgrad_tc_layer(c.p, l)) continue;
    int rc = wgrad_tc_rows(c, l, t_lo, t_hi, wg_stream);
    if (rc) return rc;
  }
  return HDPO_OK;
}
#endif

// partial slices -> gradient (and, without the overlapped form, the weight-gradient GEMMs themselves)
static int bwd_end(ChunkCtx& c) {
  const Plan& p = c.p;
  void* ws = c.ws;
  void* stream = c.stream;
  const size_t rows = static_cast<size_t>(p.T) * p.Bp;
  int splits = kSplitK;
  while (splits > 1 && (rows / splits) % BK != 0) splits >>= 1;
  const size_t rows_per = rows / splits;
  for (int l = 0; l < p.n; ++l) {
    const float* gz_hi = wsf(ws, p.o_gz[l]);
    const float* gz_lo = p.tc ? wsf(ws, p.o_gz_lo[l]) : nullptr;
    int used_splits = splits, transposed = 0, ldp = p.wp[l];
    const size_t c_slice = static_cast<size_t>(p.wp[l + 1]) * p.wp[l];
    bool done = false;
#ifndef HDPO_EMU
    if (wgrad_tc_layer(p, l)) {
      transposed = p.wp[l + 1] % 128 != 0;
      if (!p.wg_group) {  // (overlapped form: the groups already ran on the second stream)
        int rc = wgrad_tc_rows(c, l, 0, p.T - 1, stream);
        if (rc) return rc;
      }
      used_splits = p.wg_splits;
      ldp = transposed ? p.wp[l + 1] : p.wp[l];
      done = true;
    }
#endif
    if (!done) {
      GemmArgs g{};
      g.A = gz_hi;  // [rows][wp[l+1]] used transposed
      g.A2 = gz_lo;
      g.B = (l == 0) ? wsf(ws, p.o_X) : wsf(ws, p.o_act[l - 1]);  // [rows][wp[l]]  (X tape: first T blocks, full fp32)
      g.B2 = (p.tc && l > 0) ? wsf(ws, p.o_act_lo[l - 1]) : nullptr;
      g.C = wsf(ws, p.o_part_l[l]);
      g.M = p.wp[l + 1];
      g.N = p.wp[l];
      g.K = static_cast<int>(rows_per);
      g.lda = p.wp[l + 1];
      g.ldb = p.wp[l];
      g.ldc = p.wp[l];
      g.c_slice = c_slice;
      g.a_kslice = rows_per * p.wp[l + 1];
      g.b_kslice = rows_per * p.wp[l];
      int rc = sgemm<true, false, EPI_SPLITK>(g, splits, stream);
      if (rc) return rc;
    }
    const float* part = wsf(ws, p.o_part_l[l]);
    const int n_chunks = 128;
    auto k1 = colsum_stage1_kernel;
    if (p.tc && l + 1 < p.n) {  // the dgrad epilogue already reduced every 32-row block (full fp32 values)
      HDPO_LAUNCH_PDL(k1, dim3(ceil_div(p.wp[l + 1], 64), n_chunks), 256, 0, stream,
                      static_cast<const float*>(wsf(ws, p.o_csum[l])), static_cast<const float*>(nullptr), rows / 32,
                      p.wp[l + 1], n_chunks, wsf(ws, p.o_bpart));
    } else {
      HDPO_LAUNCH_PDL(k1, dim3(ceil_div(p.wp[l + 1], 64), n_chunks), 256, 0, stream, gz_hi, gz_lo, rows, p.wp[l + 1],
                      n_chunks, wsf(ws, p.o_bpart));
    }
    HDPO_LAUNCH_OK();
    auto k2 = unpack_grad_kernel;
    if (p.sym && l + 1 == p.n) {
      // projection layer: rows 0.. -> context columns of the store net's first layer (+ its bias), rows 32.. -> the
      // warehouse net's
      const sym::Cfg& sc = c.sc;
      HDPO_LAUNCH_PDL(k2, ceil_div(sc.s_w[0] * sc.C, 256), 256, 0, stream, part, used_splits, c_slice, ldp, transposed,
                      sc.s_w[0], sc.C, static_cast<const float*>(wsf(ws, p.o_bpart)), n_chunks, p.wp[l + 1],
                      sc.g_s_w0 + sc.s_in, sc.g_s_b0, c.grad, 0, sc.s_ld0);
      HDPO_LAUNCH_OK();
      HDPO_LAUNCH_PDL(k2, ceil_div(sc.w_w[0] * sc.C, 256), 256, 0, stream, part, used_splits, c_slice, ldp, transposed,
                      sc.w_w[0], sc.C, static_cast<const float*>(wsf(ws, p.o_bpart)), n_chunks, p.wp[l + 1],
                      sc.g_w_w0 + sc.Lw, sc.g_w_b0, c.grad, 32, sc.w_ld0);
      HDPO_LAUNCH_OK();
      continue;
    }
    HDPO_LAUNCH_PDL(k2, ceil_div(p.w[l + 1] * p.w[l], 256), 256, 0, stream, part, used_splits, c_slice, ldp, transposed,
                    p.w[l + 1], p.w[l], static_cast<const float*>(wsf(ws, p.o_bpart)), n_chunks, p.wp[l + 1], p.gw[l],
                    p.gb[l], c.grad, 0, p.w[l]);
    HDPO_LAUNCH_OK();
  }
  if (p.sym) return sym::reduce_slabs(c.sc, wsf(ws, p.o_slab), c.grad, stream);
  return HDPO_OK;
}

// grad[i] += extra[0][i] + extra[1][i] + ...  (fixed order: the chunked gradient is deterministic)
__global__ void __launch_bounds__(256) sum_chunk_grads_kernel(float* __restrict__ grad, const float* __restrict__ extra,
                                                              int n_extra, int P) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float s = grad[i];
  for (int c = 0; c < n_extra; ++c) s += extra[static_cast<size_t>(c) * P + i];
  grad[i] = s;
}

int backward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st, float g_total,
             float g_report, float* grad_params, void* ws, size_t ws_bytes, void* stream) {
  const Chunking ck = make_chunking(d);
  HDPO_REQUIRE(ws != nullptr && d->save_for_backward,
               "backward needs the workspace of a forward run with save_for_backward = 1");
  if (ws_bytes < ck.total) {
    set_error("workspace too small: %zu < %zu", ws_bytes, ck.total);
    return HDPO_E_WORKSPACE;
  }
  StreamFork fork;
  // overlapped weight gradients: only when EVERY chunk can cut its rows into whole-slice groups (the chunks may differ
  // in their padded row count), and by default only for a single chunk (see wg_overlap_mode)
  bool overlap = true;
  for (int i = 0; i < ck.n && overlap; ++i) overlap = make_plan(d, ck.rows[i]).wg_group > 0;
  int rc = fork.begin(ck.n, stream, overlap);
  if (rc) return rc;
  ChunkCtx ctx[kMaxChunks];
  float* extra = reinterpret_cast<float*>(static_cast<char*>(ws) + ck.grad_off);
  for (int i = 0; i < ck.n; ++i) {
    bind_chunk(&ctx[i], d, ck, i, demands, st, ws, fork.stream_of(i));
    ctx[i].grad = i == 0 ? grad_params : extra + static_cast<size_t>(i - 1) * ck.P;
    ctx[i].params = params;
    if (!overlap) ctx[i].p.wg_group = 0;  // bwd_end then runs the weight-gradient GEMMs itself
    if ((rc = bwd_begin(ctx[i]))) return rc;
  }
#ifndef HDPO_EMU
  if (ctx[0].p.persist && persist_bwd_enabled()) {
    wp::Ctx pc;
    fill_persist_ctx(&pc, ctx[0], d);
    if ((rc = wp::backward(pc, g_total, g_report))) return rc;
  } else
#endif
  for (int t = d->T - 1; t >= 0; --t) {
    const float rb = g_total + (t >= d->ignore_periods ? g_report : 0.f);
    for (int i = 0; i < ck.n; ++i)
      if ((rc = bwd_period(ctx[i], d, t, rb))) return rc;
#ifndef HDPO_EMU
    if (overlap) {
      // the sweep runs backwards in time: once period t is done the gz rows of [t, t_hi] exist, and their weight
      // gradients go to the chunk's low-priority stream, filling the SMs the dependent chain leaves idle
      const int G = ctx[0].p.wg_group;
      const int done = d->T - t;  // periods swept so far
      if (done % G == 0 || t == 0) {
        const int t_hi = t + ((done % G == 0) ? G : done % G) - 1;
        for (int i = 0; i < ck.n; ++i) {
          if ((rc = fork.rows_ready(i))) return rc;
          if ((rc = wgrad_group(ctx[i], t, t_hi, fork.wg_stream_of(i)))) return rc;
        }
      }
    }
#endif
  }
#ifndef HDPO_EMU
  if (overlap)
    for (int i = 0; i < ck.n; ++i)
      if ((rc = fork.wg_join(i))) return rc;
#endif
  for (int i = 0; i < ck.n; ++i)
    if ((rc = bwd_end(ctx[i]))) return rc;
  if ((rc = fork.end())) return rc;
  if (ck.n > 1) {
    auto k = sum_chunk_grads_kernel;
    HDPO_LAUNCH_PDL(k, ceil_div(ck.P, 256), 256, 0, stream, grad_params, static_cast<const float*>(extra), ck.n - 1, ck.P);
    HDPO_LAUNCH_OK();
  }
  return HDPO_OK;
}

}  // namespace wide
}  // namespace hdpo
