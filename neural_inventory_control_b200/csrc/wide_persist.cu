// wide_persist.cu - persistent forward / adjoint sweeps of the wide-net rollout (see wide_persist.cuh).
//
// Replaces, for a whole direction of trainer.py:181-216 (+ its autograd), the chain of ~20 dependent launches per
// period of rollout_wide.cu by ONE launch. Layout of a CTA (448 threads, one CTA per SM, clusters of 2 = CTA pairs):
//
//   warp 0      TMA producer: walks this pair's tile list, waits for the tile's input flag (release/acquire counters in
//               global memory), streams K blocks of A (128 scenario rows of this CTA, hi + lo) and of B (this CTA's half
//               of the tile's weight rows, hi + lo) into a 3-stage ring (cp.async.bulk.tensor, 128-byte swizzle).
//   warp 1      MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::tf32, M = 256, N = 128 | 64, K = 8.
//               3xTF32: hi*hi goes to a ROTATING accumulator that is closed every `kseg` K blocks (16 truncating
//               adds at kseg = 4), lo*hi + hi*lo to a per-tile cross accumulator. TMEM: [cross 0 | cross 1 | hi 0 | hi 1]
//               x 128 columns, so the MMAs of the next segment / next tile overlap the epilogue of this one.
//   warps 2-9   epilogue: tcgen05.ld every closed segment and ADD it into registers with round-to-nearest (the tensor
//               core's own accumulation truncates; short segments summed in registers keep the product fp32-grade),
//               then bias + activation + (hi, lo) split -> swizzled staging -> TMA store; or x act'(saved output)
//               (+ bias-gradient column sums) for the adjoint; finally a release-increment of the tile's output flag.
//   warps 10-13 policy head + simulator period (forward: neural_networks.py:369-427 + environment.py:110-270; adjoint:
//               their reverse), ONE THREAD per scenario over the transposed state X^T[column][scenario] so that every
//               global access of a warp is one coalesced line; row-major tapes for the GEMMs are produced through a
//               32 x 33 shared-memory transpose tile. Runs asynchronously to the GEMM pipeline of the same CTA.
//
// The tile lists are generated on the device (build_tasks_kernel) from a static schedule: pairs are grouped, a group
// serves `k` row tiles ("chains", skewed against each other by `skew` layer steps so that one chain's serial
// out-layer -> head -> first-layer section overlaps the other chain's hidden layers); all dependencies point backwards
// in the common time order, so the in-order lists cannot deadlock as long as all CTAs are co-resident (grid <= SMs).
#include "wide_persist.cuh"

#ifndef HDPO_EMU
#include <cstdlib>

#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace hdpo {
namespace wp {
using namespace tc;

constexpr int kStages = 3;
constexpr int kBK = 32;
constexpr int kABytes = 128 * kBK * 4;               // one half (hi or lo) of this CTA's A rows
constexpr int kBBytesMax = 64 * kBK * 4;             // one half of this CTA's B rows (BN = 128: 64 rows)
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytesMax;
constexpr int kRing = kStages * kStageBytes;
constexpr int kEpiWarps = 8;
constexpr int kWarps = 12;      // 384 threads: 168 registers per thread (448 threads would cap at 128)
constexpr int kHeadWarps = kWarps;  // a head CTA is all head warps (GEMM CTAs: producer + MMA + 8 epilogue warps, 2 idle)
constexpr int kBoxBytes = 32 * 128;
constexpr int kEpiStage = kEpiWarps * 2 * kBoxBytes;
constexpr int kBiasBytes = kEpiWarps * 64 * 4;  // per epilogue warp: the bias of its 64 tile columns
constexpr int kBarBytes = 256;
constexpr int kSmemTotal = kRing + kEpiStage + kBiasBytes + kBarBytes + 1024;
constexpr int kHeadBytes = kSmemTotal - 1024;  // head CTAs use the whole allocation for the staged rows of their warps
constexpr int kThreads = 32 * kWarps;
constexpr int kTmemCols = 512;
constexpr int kMaxPairs = 64;      // GEMM CTA pairs; the other SMs run the policy-head CTAs
constexpr int kMinHeadPairs = 4;
constexpr int kServerWarps = 8;     // a head task = 256 rows = 8 warps x 32 rows: a head CTA pair hosts 3 independent servers
constexpr int kServersPerPair = 2 * kWarps / kServerWarps;

enum { EPI_FWD_HIDDEN = 0, EPI_FWD_OUT = 1, EPI_DGRAD_HIDDEN = 2, EPI_DGRAD_GX = 3 };

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// ------------------------------------------------------------------------------------------------------------
// static schedule (shared by the device-side list builder and the host-side sizing)
// ------------------------------------------------------------------------------------------------------------
struct Sched {
  int T, nsteps, R, k, g, groups, skew, bwd;
  int C[HDPO_MAX_LAYERS];      // column tiles of step i
  int bn[HDPO_MAX_LAYERS];     // tile width of step i
  int nkb[HDPO_MAX_LAYERS];    // K blocks of step i
  int layer[HDPO_MAX_LAYERS];  // layer descriptor index of step i
  int epi[HDPO_MAX_LAYERS];
  int max_tasks;               // list stride per pair (gemm list and head list)
};
__host__ __device__ inline int flag_index(const Sched& s, int tt, int i, int rt) { return (tt * (s.nsteps + 1) + i) * s.R + rt; }
// pair (member q of its group) that runs column tile cc of chain c in step i of sweep position tt
__host__ __device__ inline int tile_owner(const Sched& s, int c, int i, int cc, int tt) {
  if (s.C[i] >= s.g) return (c * s.C[i] + cc) % s.g;
  const int spread = s.g / s.k > 0 ? s.g / s.k : 1;
  return (cc + c * spread + tt) % s.g;
}

struct Task {
  int4 a, b;
};

// ------------------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------------------
struct LayerDesc {
  CUtensorMap a_hi, a_lo, b_hi, b_lo, c_hi, c_lo, x_hi, x_lo;
  const float* bias;   // forward
  float* colsum;       // adjoint, hidden: [rows / 32][ldc] column sums of 32-row blocks of the output
  int act, bn, ldc;
  int a_tmul, c_tmul, x_tmul;  // tape rows per period (Bp when the tape has one slot per period, else 0)
};

struct HeadP {
  int B, Bp, S, W, L, Lw, ldx, ldy, nS, T;
  int lost, profit, has_edge, transshipment, discrete, save, ignore_periods, B_total;
  float wub, g_total, g_report;
  const int32_t* adjacency;
  const float *dT, *hT, *pT, *ltT, *whT;  // [T][S][Bp], [S][Bp], [S][Bp], [S*W][Bp], [3][W][Bp] (holding, lead, edge)
  float *XT, *OUTT;                       // [slots][ldx][Bp], [slots][ldy][Bp] transposed state / policy output
  float *X, *X_hi, *X_lo;                 // row-major state tapes (GEMM operands)
  float *gXT, *gYT;                       // adjoint: [ldx][Bp] state adjoint, [ldy][Bp] scratch
  float *gY_hi, *gY_lo;                   // adjoint: row-major tape of the output-layer adjoint (hi, lo)
  float *cost_b, *report_b, *reward_tb;
};

struct Params {
  LayerDesc L[HDPO_MAX_LAYERS];
  HeadP head;
  const Task* tasks;
  const Task* head_tasks;
  int* flags;
  int max_tasks, n_pass, kseg;
  int n_pairs;           // GEMM CTA pairs = blocks [0, 2 * n_pairs); the blocks after them are policy-head CTAs
  unsigned long long* trace;  // optional: [pair][role 0..3][trace_cap] x {tag, globaltimer ns}
  int trace_cap;
};

// ------------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kTimeoutNs = 4000000000ull;  // a wait longer than this is a protocol bug: trap, do not hang

__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void wait_timeout_trap(int what) {
  printf("[hdpo wide_persist] wait timed out (what=%d block=%d thread=%d)\n", what, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity, int what) {
  if (mbar_try(bar, parity)) return;
  const unsigned long long t0 = gtime();
  while (!mbar_try(bar, parity))
    if (gtime() - t0 > kTimeoutNs) wait_timeout_trap(what);
}
__device__ __forceinline__ void bar_wait_cluster(uint64_t* bar, uint32_t parity, int what) {
  if (mbar_try_cluster(bar, parity)) return;
  const unsigned long long t0 = gtime();
  while (!mbar_try_cluster(bar, parity))
    if (gtime() - t0 > kTimeoutNs) wait_timeout_trap(what);
}
// arrive on a barrier of the pair's leader CTA (shared::cluster address)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all lanes poll (one broadcast transaction), so every lane has acquire semantics for what follows
__device__ __forceinline__ void wait_flag(const int* f, int count, int what) {
  if (ld_acquire_gpu(f) >= count) return;
  const unsigned long long t0 = gtime();
  while (ld_acquire_gpu(f) < count) {
    __nanosleep(40);
    if (gtime() - t0 > kTimeoutNs) wait_timeout_trap(what);
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Activations other than ELU are cold paths: out of line, so that the six inlined variants (tanhf, log1pf, ...) times the
// unrolled chunks do not bloat the epilogue loop (instruction cache).
__device__ __noinline__ void act_cold_fwd(int act, float (&v)[16]) {
  dispatch_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = act_fwd_t<ACT>(v[i]);
  });
}
__device__ __noinline__ void act_cold_bwd(int act, float (&v)[16], const float (&hv)[16]) {
  dispatch_act(act, [&](auto tag) {
    constexpr int ACT = decltype(tag)::value;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= act_grad_out_t<ACT>(hv[i]);
  });
}

// 16 accumulator columns per call; `acc` indices are compile-time after unrolling
template <int NC>
__device__ __forceinline__ void tmem_accumulate(uint32_t taddr, float (&acc)[64]) {
#pragma unroll
  for (int c = 0; c < NC; c += 32) {
    uint32_t r0[16], r1[16];
    tmem_ld16_async(taddr + c, r0);
    tmem_ld16_async(taddr + c + 16, r1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      acc[c + j] += __uint_as_float(r0[j]);
      acc[c + 16 + j] += __uint_as_float(r1[j]);
    }
  }
}

// debug trace: role 0 producer, 1 MMA issuer, 2 first epilogue warp, 3 first head warp (leader CTA only)
struct Tracer {
  unsigned long long* at;
  int n, cap;
  __device__ __forceinline__ void init(const Params& p, int pair, int role) {
    cap = p.trace ? p.trace_cap : 0;
    at = p.trace ? p.trace + (static_cast<size_t>(pair) * 4 + role) * p.trace_cap * 2 : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void mark(unsigned tag) {
    if (n < cap) {
      at[2 * n] = tag;
      at[2 * n + 1] = gtime();
      ++n;
    }
  }
};
// ---- policy head + simulator period: ONE THREAD PER SCENARIO (lane = row, 32 rows per warp task) -------------------
// neural_networks.py:369-427 + environment.py:110-270 (forward) and their reverse (adjoint). Everything a thread
// touches is a column of a TRANSPOSED array ([column][scenario]: state X^T, policy output Y^T, demand, cost
// coefficients, lead times), so a warp access is one coalesced 128-byte line, there are no cross-lane reductions and
// all 32 lanes do useful work (a lane = store mapping costs ~15x the instructions per scenario: shuffles, idle lanes).
// Columns are staged global -> shared memory with cp.async in blocks of kSB stores, double-buffered, so no thread ever
// waits on a global load in its dependent chain; results leave through coalesced column stores, and the row-major
// tapes the GEMMs read (next state + its tf32 (hi, lo) split; adjoint: the output-layer adjoint) through a 32 x 33
// shared-memory transpose tile. No local memory (parameters come from the constant bank, arrays are statically
// indexed): with 227 KB of shared memory carved out there is no L1 behind it.
constexpr int kSB = 2;  // stores per staged block

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// columns col0, col0 + stride, ... (ncols of them) of a [column][Bp] array, rows [row0, row0 + 32) -> dst[j * 32 + lane]
__device__ __forceinline__ void stage_cols(float* dst, const float* src, int Bp, int row0, int col0, int stride, int ncols,
                                           int lane) {
  for (int c = lane; c < ncols * 8; c += 32) {
    const int j = c >> 3, part = c & 7;
    cp_async16(dst + j * 32 + part * 4, src + static_cast<size_t>(col0 + j * stride) * Bp + row0 + part * 4);
  }
}

__device__ __forceinline__ void store_split(float* x, float* hi, float* lo, size_t at, float v) {
  if (x) x[at] = v;
  if (hi) {
    const float a = tf32_hi(v);
    hi[at] = a;
    lo[at] = tf32_hi(v - a);
  }
}

struct TileOut {  // row-major output of per-thread column streams through a 32 x 33 shared-memory transpose tile
  float* tile;
  float *d0, *d1, *d2;  // d0: fp32 (optional), d1 / d2: tf32 (hi, lo) split (optional); rows `ld` floats apart
  int ld, lane, c;
  __device__ __forceinline__ void flush() {
    __syncwarp();
    const int col = (c - 1) / 32 * 32 + lane;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) store_split(d0, d1, d2, static_cast<size_t>(r) * ld + col, tile[r * 33 + lane]);
    __syncwarp();
  }
  __device__ __forceinline__ void emit(float v) {
    tile[lane * 33 + (c & 31)] = v;
    ++c;
    if ((c & 31) == 0) flush();
  }
};

// shared memory of one head warp (floats): [ys: S columns][2 staged blocks][32 x 33 tile]
template <bool BWD>
__host__ __device__ inline int block_cols(int W, int L) {  // columns of one staged block
  return BWD ? kSB * (1 + 3 + 2 * W + L + W) : kSB * (L + 3 + 2 * W);
}
template <bool BWD>
__host__ __device__ inline int head_warp_floats(int S, int W, int L, int Lw) {
  return 32 * (S + 4 * W + 2 * W * Lw) + 2 * 32 * block_cols<BWD>(W, L) + 32 * 33;
}

// softmax statistics of warehouse w over its connected stores (+ the constant hold logit 1.0), from the staged
// column block ys[s * 32 + lane]
__device__ __forceinline__ void softmax_stats(const HeadP& h, const float* ys, const int* adj, int w, int lane, float& mx,
                                              float& inv) {
  float m = h.transshipment ? -INFINITY : 1.f;
  for (int s = 0; s < h.S; ++s)
    if (!adj || adj[w * h.S + s]) m = fmaxf(m, ys[s * 32 + lane]);
  float sum = 0.f;
  for (int s = 0; s < h.S; ++s)
    if (!adj || adj[w * h.S + s]) sum += expf(ys[s * 32 + lane] - m);
  if (!h.transshipment) sum += expf(1.f - m);
  mx = m;
  inv = 1.f / sum;
}

__device__ __forceinline__ void head_fwd_rows(const HeadP& h, int t, int row0, int lane, float* wsm, const int* adj) {
  const int b = row0 + lane;
  const bool valid = b < h.B;
  const int Bp = h.Bp, S = h.S, W = h.W, L = h.L, Lw = h.Lw;
  const size_t Bps = static_cast<size_t>(Bp);
  const int xs = h.save ? t : (t & 1), xn = h.save ? t + 1 : ((t + 1) & 1);
  const float* XT = h.XT + static_cast<size_t>(xs) * h.ldx * Bps;
  float* XTn = h.XT + static_cast<size_t>(xn) * h.ldx * Bps + b;
  const float* YT = h.OUTT + static_cast<size_t>(h.save ? t : 0) * h.ldy * Bps;
  const float* dT = h.dT + static_cast<size_t>(t) * S * Bps;
  const size_t xstride = Bps * h.ldx;
  float* ys = wsm;                            // [S] policy-output columns of one warehouse
  float* whs = ys + 32 * S;                   // [W * Lw state | W outputs | 3 W statics] columns
  float* blk = whs + 32 * (4 * W + 2 * W * Lw);
  const int ncb = block_cols<false>(W, L);
  TileOut out;
  out.tile = blk + 2 * 32 * ncb;
  out.ld = h.ldx;
  out.lane = lane;
  out.c = 0;
  out.d0 = h.X + xn * xstride + static_cast<size_t>(row0) * h.ldx;
  const bool split = t + 1 < h.T;
  const size_t hs = h.save ? static_cast<size_t>(t + 1) : 0;
  out.d1 = split ? h.X_hi + hs * xstride + static_cast<size_t>(row0) * h.ldx : nullptr;
  out.d2 = split ? h.X_lo + hs * xstride + static_cast<size_t>(row0) * h.ldx : nullptr;

  // block s0: [X: kSB * L][d][h][p: kSB each][y: kSB * W][lt: kSB * W]
  auto stage_block = [&](int s0, float* dst) {
    const int ns = S - s0 < kSB ? S - s0 : kSB;
    stage_cols(dst, XT, Bp, row0, s0 * L, 1, ns * L, lane);
    stage_cols(dst + 32 * kSB * L, dT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * (L + 1), h.hT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * (L + 2), h.pT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * (L + 3), YT, Bp, row0, s0 * W, 1, ns * W, lane);
    stage_cols(dst + 32 * kSB * (L + 3 + W), h.ltT, Bp, row0, s0 * W, 1, ns * W, lane);
    cp_async_commit();
  };
  // ---- warehouse columns, then per warehouse the policy-output columns of its stores -> softmax statistics
  stage_cols(whs, XT, Bp, row0, h.nS, 1, W * Lw, lane);
  stage_cols(whs + 32 * W * Lw, YT, Bp, row0, S * W, 1, W, lane);
  stage_cols(whs + 32 * (W * Lw + W), h.whT, Bp, row0, 0, 1, 3 * W, lane);
  float mx[kMaxW], inv[kMaxW], W0[kMaxW], drawn[kMaxW];
#pragma unroll
  for (int w = 0; w < kMaxW; ++w) {
    mx[w] = inv[w] = W0[w] = drawn[w] = 0.f;
    if (w < W) {
      stage_cols(ys, YT, Bp, row0, w, W, S, lane);
      cp_async_commit();
      if (w + 1 == W) stage_block(0, blk);  // the first store block travels while the statistics are computed
      if (w + 1 == W) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncwarp();
      softmax_stats(h, ys, adj, w, lane, mx[w], inv[w]);
      W0[w] = whs[(w * Lw) * 32 + lane];
      __syncwarp();
    }
  }
  // ---- stores, kSB at a time
  float cost = 0.f;
  const int n_blocks = (S + kSB - 1) / kSB;
#pragma unroll 1
  for (int ib = 0; ib < n_blocks; ++ib) {
    const float* cur = blk + (ib & 1) * 32 * ncb;
    if (ib + 1 < n_blocks) {
      stage_block((ib + 1) * kSB, blk + ((ib + 1) & 1) * 32 * ncb);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kSB; ++j) {
      const int s = ib * kSB + j;
      if (s < S) {
        const float* bx = cur + (j * L) * 32 + lane;
        const float on_hand = bx[0];
        const float d = cur[(kSB * L + j) * 32 + lane], hh = cur[(kSB * (L + 1) + j) * 32 + lane],
                    pp = cur[(kSB * (L + 2) + j) * 32 + lane];
        const float raw = on_hand - d;
        cost += h.profit ? (-pp * fminf(on_hand, d) + hh * relu0(raw)) : (pp * relu0(-raw) + hh * relu0(raw));
        const float post = h.lost ? relu0(raw) : raw;
        float al[kMaxW];
        int slot[kMaxW];
#pragma unroll
        for (int w = 0; w < kMaxW; ++w) {
          al[w] = 0.f;
          slot[w] = -1;
          if (w < W && (!adj || adj[w * S + s])) {
            float a = expf(cur[(kSB * (L + 3) + j * W + w) * 32 + lane] - mx[w]) * inv[w] * W0[w];
            if (h.discrete) a = rintf(a);
            drawn[w] += a;
            al[w] = a;
            if (a != 0.f) slot[w] = static_cast<int>(cur[(kSB * (L + 3 + W) + j * W + w) * 32 + lane]) - 1;
          }
        }
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
          const float nx = k < L - 1 ? bx[(k + 1) * 32] : 0.f;
          float v = k == 0 ? post + nx : nx;
#pragma unroll
          for (int w = 0; w < kMaxW; ++w)
            if (slot[w] == k) v += al[w];
          v = valid ? v : 0.f;
          __stcg(XTn + static_cast<size_t>(s * L + k) * Bps, v);
          out.emit(v);
        }
      }
    }
    __syncwarp();  // every lane is done with `cur` before the block after next is staged into it
  }
  // ---- warehouses
#pragma unroll
  for (int w = 0; w < kMaxW; ++w) {
    if (w < W) {
      const float* xw = whs + (w * Lw) * 32 + lane;
      const float raw = W0[w] - drawn[w];
      float aw = sigmoid_f(whs[(W * Lw + w) * 32 + lane]) * h.wub;
      if (h.discrete) aw = rintf(aw);
      const float* st = whs + (W * Lw + W) * 32 + lane;  // [holding | lead | edge][W]
      float cw = st[w * 32] * relu0(raw);
      if (h.has_edge) cw += st[(2 * W + w) * 32] * aw;
      cost += cw;
      const int slotw = aw != 0.f ? static_cast<int>(st[(W + w) * 32]) - 1 : -1;
#pragma unroll 1
      for (int k = 0; k < Lw; ++k) {
        const float nx = k < Lw - 1 ? xw[(k + 1) * 32] : 0.f;
        float v = k == 0 ? raw + nx : nx;
        if (slotw == k) v += aw;
        v = valid ? v : 0.f;
        __stcg(XTn + static_cast<size_t>(h.nS + w * Lw + k) * Bps, v);
        out.emit(v);
      }
    }
  }
  for (int c = h.nS + W * Lw; c < h.ldx; ++c) {
    __stcg(XTn + static_cast<size_t>(c) * Bps, 0.f);
    out.emit(0.f);
  }
  if (valid) {
    // one addition per (scenario, period), periods ordered by the dependency chain: deterministic; RED = no round trip
    atomicAdd(h.cost_b + b, cost);
    if (h.report_b && t >= h.ignore_periods) atomicAdd(h.report_b + b, cost);
    if (h.reward_tb) h.reward_tb[static_cast<size_t>(t) * h.B_total + b] = cost;
  }
}

// Adjoint of head + period. gX^T holds the adjoint wrt X_{t+1} on entry and the direct part of the adjoint wrt X_t on
// exit (the first-layer dgrad tiles add the rest); gy goes to the row-major (hi, lo) tape of the output layer.
__device__ __forceinline__ void head_bwd_rows(const HeadP& h, int t, int row0, int lane, float* wsm, const int* adj) {
  const int b = row0 + lane;
  const bool valid = b < h.B;
  const int Bp = h.Bp, S = h.S, W = h.W, L = h.L, Lw = h.Lw;
  const size_t Bps = static_cast<size_t>(Bp);
  const float* XT = h.XT + static_cast<size_t>(t) * h.ldx * Bps;
  const float* YT = h.OUTT + static_cast<size_t>(t) * h.ldy * Bps;
  const float* dT = h.dT + static_cast<size_t>(t) * S * Bps;
  float* G = h.gXT + b;
  float* GY = h.gYT + b;
  const float rb = h.g_total + (t >= h.ignore_periods ? h.g_report : 0.f);
  const size_t ystride = Bps * h.ldy;
  float* ys = wsm;
  float* whs = ys + 32 * S;  // [W * Lw adjoint | W on-hand | W outputs | 3 W statics] columns
  float* blk = whs + 32 * (4 * W + 2 * W * Lw);
  const int ncb = block_cols<true>(W, L);
  TileOut out;
  out.tile = blk + 2 * 32 * ncb;
  out.ld = h.ldy;
  out.lane = lane;
  out.c = 0;
  out.d0 = nullptr;
  out.d1 = h.gY_hi + static_cast<size_t>(t) * ystride + static_cast<size_t>(row0) * h.ldy;
  out.d2 = h.gY_lo + static_cast<size_t>(t) * ystride + static_cast<size_t>(row0) * h.ldy;
  const int SW = S * W;

  // pass-1 block s0: [x0: kSB][d][h][p: kSB each][y: kSB * W][lt: kSB * W][g: kSB * L]
  auto stage_block1 = [&](int s0, float* dst) {
    const int ns = S - s0 < kSB ? S - s0 : kSB;
    stage_cols(dst, XT, Bp, row0, s0 * L, L, ns, lane);
    stage_cols(dst + 32 * kSB, dT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * 2, h.hT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * 3, h.pT, Bp, row0, s0, 1, ns, lane);
    stage_cols(dst + 32 * kSB * 4, YT, Bp, row0, s0 * W, 1, ns * W, lane);
    stage_cols(dst + 32 * kSB * (4 + W), h.ltT, Bp, row0, s0 * W, 1, ns * W, lane);
    stage_cols(dst + 32 * kSB * (4 + 2 * W), h.gXT, Bp, row0, s0 * L, 1, ns * L, lane);
    cp_async_commit();
  };
  // pass-2 block s0: [y: kSB * W][galloc: kSB * W]
  auto stage_block2 = [&](int s0, float* dst) {
    const int ns = S - s0 < kSB ? S - s0 : kSB;
    stage_cols(dst, YT, Bp, row0, s0 * W, 1, ns * W, lane);
    stage_cols(dst + 32 * kSB * W, h.gYT, Bp, row0, s0 * W, 1, ns * W, lane);
    cp_async_commit();
  };
  stage_cols(whs, h.gXT, Bp, row0, h.nS, 1, W * Lw, lane);
  stage_cols(whs + 32 * W * Lw, XT, Bp, row0, h.nS, Lw, W, lane);
  stage_cols(whs + 32 * (W * Lw + W), YT, Bp, row0, SW, 1, W, lane);
  stage_cols(whs + 32 * (W * Lw + 2 * W), h.whT, Bp, row0, 0, 1, 3 * W, lane);
  float mx[kMaxW], inv[kMaxW], W0[kMaxW], graw[kMaxW], dot[kMaxW], gyw[kMaxW];
#pragma unroll
  for (int w = 0; w < kMaxW; ++w) {
    mx[w] = inv[w] = W0[w] = graw[w] = dot[w] = gyw[w] = 0.f;
    if (w < W) {
      stage_cols(ys, YT, Bp, row0, w, W, S, lane);
      cp_async_commit();
      if (w + 1 == W) stage_block1(0, blk);
      if (w + 1 == W) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncwarp();
      softmax_stats(h, ys, adj, w, lane, mx[w], inv[w]);
      W0[w] = whs[(W * Lw + w) * 32 + lane];
      float drawn = 0.f;
      for (int s = 0; s < S; ++s)
        if (!adj || adj[w * S + s]) drawn += expf(ys[s * 32 + lane] - mx[w]) * inv[w] * W0[w];
      // warehouse w: g_raw_w feeds the allocation adjoints of the stores
      const float* st = whs + (W * Lw + 2 * W) * 32 + lane;  // [holding | lead | edge][W]
      const float* gw = whs + (w * Lw) * 32 + lane;
      float* gwo = G + static_cast<size_t>(h.nS + w * Lw) * Bps;
      const float raw = W0[w] - drawn;
      const float sg = sigmoid_f(whs[(W * Lw + W + w) * 32 + lane]);
      const float aw = sg * h.wub;
      const int slotw = aw != 0.f ? static_cast<int>(st[(W + w) * 32]) - 1 : -1;
      float gaw = 0.f;
#pragma unroll 1
      for (int k = 0; k < Lw; ++k) {
        const float cur = gw[k * 32];
        if (slotw == k) gaw = cur;
        if (k + 1 < Lw) __stcg(gwo + static_cast<size_t>(k + 1) * Bps, valid ? cur : 0.f);  // new gw[k + 1] = old gw[k]
      }
      if (h.has_edge) gaw += rb * st[(2 * W + w) * 32];
      graw[w] = rb * st[w * 32] * ge0(raw) + gw[0];
      gyw[w] = gaw * h.wub * sg * (1.f - sg);
      __syncwarp();
    }
  }
  // ---- pass 1 over the stores: dynamics adjoint, allocation adjoints (parked in gY^T), softmax inner products
  const int n_blocks = (S + kSB - 1) / kSB;
#pragma unroll 1
  for (int ib = 0; ib < n_blocks; ++ib) {
    const float* cur = blk + (ib & 1) * 32 * ncb;
    if (ib + 1 < n_blocks) {
      stage_block1((ib + 1) * kSB, blk + ((ib + 1) & 1) * 32 * ncb);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kSB; ++j) {
      const int s = ib * kSB + j;
      if (s < S) {
        const float on_hand = cur[j * 32 + lane];
        const float d = cur[(kSB + j) * 32 + lane], hh = cur[(kSB * 2 + j) * 32 + lane], pp = cur[(kSB * 3 + j) * 32 + lane];
        const float* gs = cur + (kSB * (4 + 2 * W) + j * L) * 32 + lane;
        float* gso = G + static_cast<size_t>(s * L) * Bps;
        const float raw = on_hand - d;
        float share[kMaxW], ga[kMaxW];
        int slot[kMaxW];
#pragma unroll
        for (int w = 0; w < kMaxW; ++w) {
          share[w] = ga[w] = 0.f;
          slot[w] = -1;
          if (w < W && (!adj || adj[w * S + s])) {
            share[w] = expf(cur[(kSB * 4 + j * W + w) * 32 + lane] - mx[w]) * inv[w];
            if (share[w] * W0[w] != 0.f) slot[w] = static_cast<int>(cur[(kSB * (4 + W) + j * W + w) * 32 + lane]) - 1;
          }
        }
        const float gn0 = gs[0];
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
          const float c = gs[k * 32];
#pragma unroll
          for (int w = 0; w < kMaxW; ++w)
            if (slot[w] == k) ga[w] = c;
          if (k + 1 < L) __stcg(gso + static_cast<size_t>(k + 1) * Bps, valid ? c : 0.f);  // new g[k + 1] = old g[k]
        }
        float g0;
        if (h.profit) {
          const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
          g0 = rb * (-pp * tie + hh * ge0(raw));
        } else {
          g0 = rb * (-pp * le0(raw) + hh * ge0(raw));
        }
        g0 += h.lost ? gn0 * ge0(raw) : gn0;
        __stcg(gso, valid ? g0 : 0.f);
#pragma unroll
        for (int w = 0; w < kMaxW; ++w) {
          if (w < W) {
            const float galloc = ga[w] - graw[w];
            dot[w] += galloc * share[w];
            __stcg(GY + static_cast<size_t>(s * W + w) * Bps, galloc);
          }
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int w = 0; w < kMaxW; ++w)
    if (w < W) __stcg(G + static_cast<size_t>(h.nS + w * Lw) * Bps, valid ? graw[w] + dot[w] : 0.f);
  // ---- pass 2: softmax backward -> gy in column order s * W + w, then the warehouse-order columns, then zero padding
  __threadfence_block();  // this thread re-reads the allocation adjoints it parked in gY^T (through cp.async)
  stage_block2(0, blk);
#pragma unroll 1
  for (int ib = 0; ib < n_blocks; ++ib) {
    const float* cur = blk + (ib & 1) * 32 * ncb;
    if (ib + 1 < n_blocks) {
      stage_block2((ib + 1) * kSB, blk + ((ib + 1) & 1) * 32 * ncb);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kSB; ++j) {
      const int s = ib * kSB + j;
      if (s < S) {
#pragma unroll
        for (int w = 0; w < kMaxW; ++w) {
          if (w < W) {
            float v = 0.f;
            if (!adj || adj[w * S + s]) {
              const float pw = expf(cur[(j * W + w) * 32 + lane] - mx[w]) * inv[w];
              v = pw * (cur[(kSB * W + j * W + w) * 32 + lane] * W0[w] - dot[w] * W0[w]);
            }
            out.emit(valid ? v : 0.f);
          }
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int w = 0; w < kMaxW; ++w)
    if (w < W) out.emit(valid ? gyw[w] : 0.f);
  for (int c = SW + W; c < h.ldy; ++c) out.emit(0.f);
}

// ------------------------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1) persist_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  unsigned char* ring = smem;
  unsigned char* epi_stage = smem + kRing;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRing + kEpiStage + kBiasBytes);
  float* bias_all = reinterpret_cast<float*>(smem + kRing + kEpiStage);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* hi_full = empty_bar + kStages;
  uint64_t* hi_empty = hi_full + 2;
  uint64_t* cr_full = hi_empty + 2;
  uint64_t* cr_empty = cr_full + 2;
  uint64_t* aux_bar = cr_empty + 2;  // one per epilogue warp (adjoint: TMA loads of the saved-activation boxes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const bool three = p.n_pass == 3;

  if (pair >= p.n_pairs) {
    // ===== policy head + simulator period CTAs: the 2 x 12 warps of a CTA pair share the rows of a head task =====
    // (own SMs: their L2 return path and instruction cache are not competing with a GEMM pipeline)
    const int hpair = pair - p.n_pairs;
    const int hw_id = static_cast<int>(rank) * kHeadWarps + warp;            // warp among the pair's 24
    const int server = hpair * kServersPerPair + hw_id / kServerWarps;         // 8 warps = one head server
    const int blk_id = hw_id % kServerWarps;                                   // its 32-row block of every task
    int* adj_s = reinterpret_cast<int*>(smem);
    const int adj_floats = p.head.adjacency ? ((p.head.S * p.head.W + 3) & ~3) : 0;
    if (p.head.adjacency) {  // adjacency masks: staged once
      for (int e = threadIdx.x; e < p.head.S * p.head.W; e += kThreads) adj_s[e] = __ldg(p.head.adjacency + e);
    }
    __syncthreads();
    float* wsm = reinterpret_cast<float*>(smem) + adj_floats + warp * ((kHeadBytes / 4 - adj_floats) / kHeadWarps / 4 * 4);
    const int* adj = p.head.adjacency ? adj_s : nullptr;
    const Task* tp = p.head_tasks + static_cast<size_t>(server) * p.max_tasks;
    Tracer tr;
    tr.init(p, hpair, 3);
    const bool tracing = rank == 0 && warp == 0 && lane == 0;
    for (;; ++tp) {
      const int4 a = __ldg(&tp->a), b = __ldg(&tp->b);
      if ((a.x & 0xff) == 0) break;
      if (tracing) tr.mark(0x100);
      if (a.w >= 0) wait_flag(p.flags + a.w, b.x, 9);
      if (tracing) tr.mark(0x200);
      const int row0 = a.z * kRowTile + blk_id * 32;
      if (BWD) head_bwd_rows(p.head, a.y, row0, lane, wsm, adj);
      else head_fwd_rows(p.head, a.y, row0, lane, wsm, adj);
      if (tracing) tr.mark(0x300);
      __threadfence();
      __syncwarp();
      if (lane == 0) red_release_gpu(p.flags + b.y, 1);
      if (tracing) tr.mark(0x400);
    }
    return;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hi_full[i], 1);
      mbar_init(&cr_full[i], 1);
      mbar_init(&hi_empty[i], 2 * kEpiWarps);
      mbar_init(&cr_empty[i], 2 * kEpiWarps);
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&aux_bar[w], 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    const Task* tp = p.tasks + static_cast<size_t>(pair) * p.max_tasks;
    uint32_t s = 0, ph = 0;
    Tracer tr;
    tr.init(p, pair, 0);
    const bool tracing = rank == 0 && lane == 0;
    for (;; ++tp) {
      const int4 a = __ldg(&tp->a), b = __ldg(&tp->b);
      if ((a.x & 0xff) == 0) break;
      const LayerDesc& L = p.L[(a.x >> 8) & 0xff];
      if (tracing) tr.mark(0x100 | ((a.x >> 8) & 0xff));  // reached the task
      if (b.y >= 0) {
        wait_flag(p.flags + b.y, b.z, 1);
        fence_proxy_async_all();
      }
      if (tracing) tr.mark(0x200 | ((a.x >> 8) & 0xff));  // inputs ready
      const int bn = L.bn;
      const uint32_t tx = (three ? 2u : 1u) * static_cast<uint32_t>(kABytes + (bn / 2) * 128) * 2u;
      const int arow = L.a_tmul * a.y + a.z * kRowTile + static_cast<int>(rank) * 128;
      const int brow = a.w + static_cast<int>(rank) * (bn / 2);
      for (int kb = 0; kb < b.x; ++kb) {
        bar_wait(&empty_bar[s], ph ^ 1, 2);
        unsigned char* st = ring + s * kStageBytes;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full_bar[s], tx);
          const uint32_t bar = mapa_u32(&full_bar[s], 0);
          tma_load_2d_pair(st, &L.a_hi, bar, kb * kBK, arow);
          tma_load_2d_pair(st + 2 * kABytes, &L.b_hi, bar, kb * kBK, brow);
          if (three) {
            tma_load_2d_pair(st + kABytes, &L.a_lo, bar, kb * kBK, arow);
            tma_load_2d_pair(st + 2 * kABytes + kBBytesMax, &L.b_lo, bar, kb * kBK, brow);
          }
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (tracing) tr.mark(0x300 | ((a.x >> 8) & 0xff));  // all loads issued
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA of the pair) =====
    if (rank == 0) {
      const Task* tp = p.tasks + static_cast<size_t>(pair) * p.max_tasks;
      uint32_t s = 0, ph = 0, hseg = 0, tcnt = 0;
      Tracer tr;
      tr.init(p, pair, 1);
      for (;; ++tp) {
        const int4 a = __ldg(&tp->a), b = __ldg(&tp->b);
        if ((a.x & 0xff) == 0) break;
        const LayerDesc& L = p.L[(a.x >> 8) & 0xff];
        const int bn = L.bn, n_kb = b.x;
        if (lane == 0) tr.mark(0x100 | ((a.x >> 8) & 0xff));
        // D = f32 (bit 4), A = B = tf32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | ((256u >> 4) << 24);
        const uint32_t cb = tcnt & 1;
        const uint32_t d_cr = tmem + cb * 128;
        if (three) {
          bar_wait_cluster(&cr_empty[cb], ((tcnt >> 1) & 1) ^ 1, 3);
          tc_fence_after();
        }
        int in_seg = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
          const uint32_t hb = hseg & 1;
          if (in_seg == 0) {
            bar_wait_cluster(&hi_empty[hb], ((hseg >> 1) & 1) ^ 1, 4);
            tc_fence_after();
          }
          bar_wait(&full_bar[s], ph, 5);
          tc_fence_after();
          if (lane == 0 && kb == 0) tr.mark(0x200 | ((a.x >> 8) & 0xff));  // first operands landed
          const bool seg_last = (in_seg == p.kseg - 1) || (kb == n_kb - 1);
          if (elect_one()) {
            const uint32_t st = smem_u32(ring + s * kStageBytes);
            const uint64_t a_hi = make_kmajor_desc(st), a_lo = make_kmajor_desc(st + kABytes);
            const uint64_t b_hi = make_kmajor_desc(st + 2 * kABytes), b_lo = make_kmajor_desc(st + 2 * kABytes + kBBytesMax);
            const uint32_t d_hi = tmem + 256 + hb * 128;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * 8 * 4) >> 4);
              umma_tf32_pair(d_hi, a_hi + adv, b_hi + adv, idesc, (in_seg == 0 && k == 0) ? 0u : 1u);
              if (three) {
                umma_tf32_pair(d_cr, a_lo + adv, b_hi + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
                umma_tf32_pair(d_cr, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            }
            umma_commit_pair(&empty_bar[s]);
            if (seg_last) umma_commit_pair(&hi_full[hb]);
            if (three && kb == n_kb - 1) umma_commit_pair(&cr_full[cb]);
          }
          __syncwarp();
          if (seg_last) {
            ++hseg;
            in_seg = 0;
          } else {
            ++in_seg;
          }
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        ++tcnt;
        if (lane == 0) tr.mark(0x300 | ((a.x >> 8) & 0xff));  // all MMAs issued
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===== epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q + 32) = rows of this CTA's half of the tile =====
    const int q = warp & 3, we = warp - 2, half = we >> 2;
    unsigned char* stg = epi_stage + we * (2 * kBoxBytes);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t hi_empty_c = mapa_u32(&hi_empty[0], 0), cr_empty_c = mapa_u32(&cr_empty[0], 0);
    const Task* tp = p.tasks + static_cast<size_t>(pair) * p.max_tasks;
    uint32_t hseg = 0, tcnt = 0, aux_cnt = 0;
    (void)aux_cnt;
    Tracer tr;
    tr.init(p, pair, 2);
    const bool tracing = rank == 0 && warp == 2 && lane == 0;
    // Deferred publish: the flag of a finished tile is incremented once its TMA stores have completed, but the warp
    // does not sit on that wait - it moves on to the next tile's accumulators and publishes when it would block anyway
    // (never later: the next tile of this pair may transitively depend on the pending flag).
    int pending = -1;
    auto publish_pending = [&]() {
      if (pending >= 0) {
        if (lane == 0) {
          tma_store_wait_all();
          fence_proxy_async_all();
          red_release_gpu(p.flags + pending, 1);
        }
        pending = -1;
      }
    };
    for (;; ++tp) {
      const int4 a = __ldg(&tp->a), b = __ldg(&tp->b);
      if ((a.x & 0xff) == 0) break;
      const LayerDesc& L = p.L[(a.x >> 8) & 0xff];
      const int epi = (a.x >> 16) & 0xff;
      if (tracing) tr.mark(0x100 | ((a.x >> 8) & 0xff));
      const int t = a.y, col0 = a.w, n_kb = b.x;
      const bool wide = L.bn == 128;  // this warp drains 64 (wide) or 32 columns
      const int cbase = half * (wide ? 64 : 32);
      const int trow = a.z * kRowTile + static_cast<int>(rank) * 128 + q * 32;  // first scenario row of this warp
      const int n0 = col0 + cbase;
      float* bias_w = bias_all + we * 64;
      // Global loads cost ~1.4 us here (the TMA operand stream keeps the SM's L2 return path full), so whatever the
      // finalize step needs is requested NOW and arrives while the tile's MMAs run.
      if (!BWD) {
        __syncwarp();
        bias_w[lane] = __ldg(L.bias + n0 + lane);
        if (wide) bias_w[lane + 32] = __ldg(L.bias + n0 + 32 + lane);
        __syncwarp();
      }
      auto issue_aux = [&](int j) {  // adjoint: box j of the tile this epilogue combines with -> staging
        if (lane == 0) {
          tma_store_wait_read();
          fence_proxy_async_all();
          mbar_expect_tx(&aux_bar[we], 2 * kBoxBytes);
          tma_load_2d(stg, &L.x_hi, &aux_bar[we], n0 + 32 * j, L.x_tmul * t + trow);
          tma_load_2d(stg + kBoxBytes, &L.x_lo, &aux_bar[we], n0 + 32 * j, L.x_tmul * t + trow);
        }
      };
      if (BWD && epi == EPI_DGRAD_HIDDEN) {
        publish_pending();  // the staging area is about to be overwritten: the pending tile's stores must have left it
        issue_aux(0);       // saved activations: complete since the forward sweep
      }
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < n_kb; kb0 += p.kseg) {
        const uint32_t hb = hseg & 1;
        if (pending >= 0 && !__all_sync(0xffffffffu, mbar_try(&hi_full[hb], (hseg >> 1) & 1))) publish_pending();
        bar_wait(&hi_full[hb], (hseg >> 1) & 1, 6);
        tc_fence_after();
        const uint32_t taddr = lane_base + 256 + hb * 128 + cbase;
        if (wide) tmem_accumulate<64>(taddr, acc);
        else tmem_accumulate<32>(taddr, acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(hi_empty_c + hb * 8);
        ++hseg;
      }
      if (three) {
        const uint32_t cb = tcnt & 1;
        if (pending >= 0 && !__all_sync(0xffffffffu, mbar_try(&cr_full[cb], (tcnt >> 1) & 1))) publish_pending();
        bar_wait(&cr_full[cb], (tcnt >> 1) & 1, 7);
        tc_fence_after();
        const uint32_t taddr = lane_base + cb * 128 + cbase;
        if (wide) tmem_accumulate<64>(taddr, acc);
        else tmem_accumulate<32>(taddr, acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cr_empty_c + cb * 8);
      }
      ++tcnt;
      if (tracing) tr.mark(0x200 | ((a.x >> 8) & 0xff));  // accumulators drained
      publish_pending();
      // ---- finalize: this warp's 32 rows x (64 | 32) columns ----
      if (!BWD && epi == EPI_FWD_HIDDEN) {
        const int grow = L.c_tmul * t + trow;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j == 0 || wide) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float v[16];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_w + 32 * j + 16 * hh + 4 * j4);
                v[4 * j4 + 0] = acc[32 * j + 16 * hh + 4 * j4 + 0] + b4.x;
                v[4 * j4 + 1] = acc[32 * j + 16 * hh + 4 * j4 + 1] + b4.y;
                v[4 * j4 + 2] = acc[32 * j + 16 * hh + 4 * j4 + 2] + b4.z;
                v[4 * j4 + 3] = acc[32 * j + 16 * hh + 4 * j4 + 3] + b4.w;
              }
              if (L.act == HDPO_ACT_ELU) {
                elu_inplace(v);
              } else {
                act_cold_fwd(L.act, v);
              }
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                float4 hi, lo;
                hi.x = tf32_hi(v[4 * j4 + 0]);
                hi.y = tf32_hi(v[4 * j4 + 1]);
                hi.z = tf32_hi(v[4 * j4 + 2]);
                hi.w = tf32_hi(v[4 * j4 + 3]);
                lo.x = tf32_hi(v[4 * j4 + 0] - hi.x);
                lo.y = tf32_hi(v[4 * j4 + 1] - hi.y);
                lo.z = tf32_hi(v[4 * j4 + 2] - hi.z);
                lo.w = tf32_hi(v[4 * j4 + 3] - hi.w);
                const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
                *reinterpret_cast<float4*>(stg + lane * 128 + chunk) = hi;
                *reinterpret_cast<float4*>(stg + kBoxBytes + lane * 128 + chunk) = lo;
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&L.c_hi, stg, n0 + 32 * j, grow);
              tma_store_2d(&L.c_lo, stg + kBoxBytes, n0 + 32 * j, grow);
              tma_store_commit();
            }
          }
        }
      } else if (!BWD && epi == EPI_FWD_OUT) {
        // output layer (64-column tiles: 32 columns per warp): fp32 row-major tape by TMA + transposed copy for the head
        const int grow = L.c_tmul * t + trow;
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        float* yt = p.head.OUTT + static_cast<size_t>(p.head.save ? t : 0) * p.head.ldy * p.head.Bp +
                    static_cast<size_t>(n0) * p.head.Bp + trow + lane;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_w + 4 * j4);
          float4 v;
          v.x = acc[4 * j4 + 0] + b4.x;
          v.y = acc[4 * j4 + 1] + b4.y;
          v.z = acc[4 * j4 + 2] + b4.z;
          v.w = acc[4 * j4 + 3] + b4.w;
          *reinterpret_cast<float4*>(stg + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = v;
          __stcg(yt + static_cast<size_t>(4 * j4 + 0) * p.head.Bp, v.x);
          __stcg(yt + static_cast<size_t>(4 * j4 + 1) * p.head.Bp, v.y);
          __stcg(yt + static_cast<size_t>(4 * j4 + 2) * p.head.Bp, v.z);
          __stcg(yt + static_cast<size_t>(4 * j4 + 3) * p.head.Bp, v.w);
        }
        __threadfence();  // the transposed copy (plain stores) must be visible before the tile's flag is published
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&L.c_hi, stg, n0, grow);
          tma_store_commit();
        }
      } else if (BWD && epi == EPI_DGRAD_HIDDEN) {
        const int grow = L.c_tmul * t + trow;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j == 0 || wide) {
            if (j == 1) issue_aux(1);
            bar_wait(&aux_bar[we], aux_cnt & 1, 8);
            ++aux_cnt;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float v[16], hv[16];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
                const float4 x0 = *reinterpret_cast<const float4*>(stg + lane * 128 + chunk);
                const float4 x1 = *reinterpret_cast<const float4*>(stg + kBoxBytes + lane * 128 + chunk);
                hv[4 * j4 + 0] = x0.x + x1.x;
                hv[4 * j4 + 1] = x0.y + x1.y;
                hv[4 * j4 + 2] = x0.z + x1.z;
                hv[4 * j4 + 3] = x0.w + x1.w;
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = acc[32 * j + 16 * hh + i];
              if (L.act == HDPO_ACT_ELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= elu_grad_from_out(hv[i]);
              } else {
                act_cold_bwd(L.act, v, hv);
              }
              if (L.colsum) {
                // bias-gradient partials: column sums over this warp's 32 rows (reduce-scatter butterfly)
                float w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = v[i];
#pragma unroll
                for (int hw = 8, o = 16; hw >= 1; hw >>= 1, o >>= 1) {
                  const bool up = (lane & o) != 0;
#pragma unroll
                  for (int i = 0; i < hw; ++i) {
                    const float send = up ? w[i] : w[i + hw];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, o);
                    w[i] = (up ? w[i + hw] : w[i]) + recv;
                  }
                }
                w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
                if ((lane & 1) == 0) {
                  const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                  L.colsum[(static_cast<size_t>(grow) >> 5) * L.ldc + n0 + 32 * j + 16 * hh + col] = w[0];
                }
              }
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                float4 hi, lo;
                hi.x = tf32_hi(v[4 * j4 + 0]);
                hi.y = tf32_hi(v[4 * j4 + 1]);
                hi.z = tf32_hi(v[4 * j4 + 2]);
                hi.w = tf32_hi(v[4 * j4 + 3]);
                lo.x = tf32_hi(v[4 * j4 + 0] - hi.x);
                lo.y = tf32_hi(v[4 * j4 + 1] - hi.y);
                lo.z = tf32_hi(v[4 * j4 + 2] - hi.z);
                lo.w = tf32_hi(v[4 * j4 + 3] - hi.w);
                const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
                *reinterpret_cast<float4*>(stg + lane * 128 + chunk) = hi;
                *reinterpret_cast<float4*>(stg + kBoxBytes + lane * 128 + chunk) = lo;
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&L.c_hi, stg, n0 + 32 * j, grow);
              tma_store_2d(&L.c_lo, stg + kBoxBytes, n0 + 32 * j, grow);
              tma_store_commit();
            }
          }
        }
      } else if (BWD && epi == EPI_DGRAD_GX) {
        // first-layer dgrad: add the product to the (transposed) state adjoint, one coalesced line per column
        float* g = p.head.gXT + static_cast<size_t>(n0) * p.head.Bp + trow + lane;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          if (j < 32 || wide) {
            float* at = g + static_cast<size_t>(j) * p.head.Bp;
            __stcg(at, __ldcg(at) + acc[j]);
          }
        }
        __threadfence();
        __syncwarp();
      }
      // publish: every store of this warp is complete and visible before the tile's flag is incremented
      if (tracing) tr.mark(0x300 | ((a.x >> 8) & 0xff));  // tile staged, stores issued
      pending = b.w;
    }
    publish_pending();
  }
  // (warps 10, 11 of a GEMM CTA are idle: the block size is what a policy-head CTA wants)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may release shared memory / TMEM the pair's MMAs and commits still use
  if (warp == 1) tmem_dealloc_pair(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------
// multi-tile layer GEMM: the pipeline of the persistent kernel (TMA ring -> tcgen05 pair MMAs -> rotating TMEM
// accumulators -> register-accumulating epilogue warps) behind the per-layer interface of gemm_tc.cuh. One launch =
// one layer of one period; every CTA pair walks the tiles pair, pair + n_pairs, ... of the launch, so the epilogue of a
// tile (tcgen05.ld, activation, (hi, lo) split, TMA stores: ~4 us) overlaps the MMAs of the pair's next tile and the
// launch set-up (barriers, TMEM, cluster sync: ~3 us) is paid once per pair instead of once per tile.
// ------------------------------------------------------------------------------------------------------------
struct MultiArgs {
  int M, N, K, bn, n_pass, kseg, n_pairs;
  int a_row0, b_row0, c_row0, x_row0, ldc, act;
  const float* bias;
  float* colsum;
};
constexpr int kMultiThreads = 32 * (2 + kEpiWarps);
constexpr int kMultiSmem = kSmemTotal;

template <int EPI>
__global__ void __launch_bounds__(kMultiThreads, 1) multi_kernel(const __grid_constant__ tc::GemmTcMaps tm, const MultiArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
  unsigned char* ring = smem;
  unsigned char* epi_stage = smem + kRing;
  float* bias_all = reinterpret_cast<float*>(smem + kRing + kEpiStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRing + kEpiStage + kBiasBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* hi_full = empty_bar + kStages;
  uint64_t* hi_empty = hi_full + 2;
  uint64_t* cr_full = hi_empty + 2;
  uint64_t* cr_empty = cr_full + 2;
  uint64_t* aux_bar = cr_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const bool three = g.n_pass == 3;
  const int bn = g.bn, n_kb = g.K / kBK;
  const int tiles_n = g.N / bn;
  const int n_tiles = (g.M / kRowTile) * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm.a_hi);
    prefetch_tmap(&tm.b_hi);
    if (three) {
      prefetch_tmap(&tm.a_lo);
      prefetch_tmap(&tm.b_lo);
    }
    prefetch_tmap(&tm.c0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hi_full[i], 1);
      mbar_init(&cr_full[i], 1);
      mbar_init(&hi_empty[i], 2 * kEpiWarps);
      mbar_init(&cr_empty[i], 2 * kEpiWarps);
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(&aux_bar[w], 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // everything above overlapped the tail of the previous kernel in the stream (PDL); from here on we read its output
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer =====
    uint32_t s = 0, ph = 0;
    const uint32_t tx = (three ? 2u : 1u) * static_cast<uint32_t>(kABytes + (bn / 2) * 128) * 2u;
    for (int tile = pair; tile < n_tiles; tile += g.n_pairs) {
      const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
      const int arow = g.a_row0 + mt * kRowTile + static_cast<int>(rank) * 128;
      const int brow = g.b_row0 + nt * bn + static_cast<int>(rank) * (bn / 2);
      for (int kb = 0; kb < n_kb; ++kb) {
        bar_wait(&empty_bar[s], ph ^ 1, 2);
        unsigned char* st = ring + s * kStageBytes;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full_bar[s], tx);
          const uint32_t bar = mapa_u32(&full_bar[s], 0);
          tma_load_2d_pair(st, &tm.a_hi, bar, kb * kBK, arow);
          tma_load_2d_pair(st + 2 * kABytes, &tm.b_hi, bar, kb * kBK, brow);
          if (three) {
            tma_load_2d_pair(st + kABytes, &tm.a_lo, bar, kb * kBK, arow);
            tma_load_2d_pair(st + 2 * kABytes + kBBytesMax, &tm.b_lo, bar, kb * kBK, brow);
          }
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA of the pair) =====
    if (rank == 0) {
      uint32_t s = 0, ph = 0, hseg = 0, tcnt = 0;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | ((256u >> 4) << 24);
      for (int tile = pair; tile < n_tiles; tile += g.n_pairs) {
        const uint32_t cb = tcnt & 1;
        const uint32_t d_cr = tmem + cb * 128;
        if (three) {
          bar_wait_cluster(&cr_empty[cb], ((tcnt >> 1) & 1) ^ 1, 3);
          tc_fence_after();
        }
        int in_seg = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
          const uint32_t hb = hseg & 1;
          if (in_seg == 0) {
            bar_wait_cluster(&hi_empty[hb], ((hseg >> 1) & 1) ^ 1, 4);
            tc_fence_after();
          }
          bar_wait(&full_bar[s], ph, 5);
          tc_fence_after();
          const bool seg_last = (in_seg == g.kseg - 1) || (kb == n_kb - 1);
          if (elect_one()) {
            const uint32_t st = smem_u32(ring + s * kStageBytes);
            const uint64_t a_hi = make_kmajor_desc(st), a_lo = make_kmajor_desc(st + kABytes);
            const uint64_t b_hi = make_kmajor_desc(st + 2 * kABytes), b_lo = make_kmajor_desc(st + 2 * kABytes + kBBytesMax);
            const uint32_t d_hi = tmem + 256 + hb * 128;
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * 8 * 4) >> 4);
              umma_tf32_pair(d_hi, a_hi + adv, b_hi + adv, idesc, (in_seg == 0 && k == 0) ? 0u : 1u);
              if (three) {
                umma_tf32_pair(d_cr, a_lo + adv, b_hi + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
                umma_tf32_pair(d_cr, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            }
            umma_commit_pair(&empty_bar[s]);
            if (seg_last) umma_commit_pair(&hi_full[hb]);
            if (three && kb == n_kb - 1) umma_commit_pair(&cr_full[cb]);
          }
          __syncwarp();
          if (seg_last) {
            ++hseg;
            in_seg = 0;
          } else {
            ++in_seg;
          }
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        ++tcnt;
      }
    }
  } else {
    // ===== epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q + 32) = rows of this CTA's half of the tile =====
    const int q = warp & 3, we = warp - 2, half = we >> 2;
    unsigned char* stg = epi_stage + we * (2 * kBoxBytes);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t hi_empty_c = mapa_u32(&hi_empty[0], 0), cr_empty_c = mapa_u32(&cr_empty[0], 0);
    const bool wide = bn == 128;  // this warp drains 64 (wide) or 32 columns
    const int cbase = half * (wide ? 64 : 32);
    float* bias_w = bias_all + we * 64;
    uint32_t hseg = 0, tcnt = 0, aux_cnt = 0;
    (void)aux_cnt;
    (void)bias_w;
    constexpr bool kAux = EPI == tc::EPI_DGRAD_HIDDEN || EPI == tc::EPI_DGRAD_ACCUM;
    constexpr bool kSplit = EPI == tc::EPI_FWD_HIDDEN || EPI == tc::EPI_DGRAD_HIDDEN;
    for (int tile = pair; tile < n_tiles; tile += g.n_pairs) {
      const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
      const int trow = mt * kRowTile + static_cast<int>(rank) * 128 + q * 32;  // first row of this warp inside the GEMM
      const int n0 = nt * bn + cbase;
      const int grow = g.c_row0 + trow, xrow = g.x_row0 + trow;
      if (EPI == tc::EPI_FWD_HIDDEN || EPI == tc::EPI_FWD_OUT) {
        __syncwarp();
        bias_w[lane] = __ldg(g.bias + n0 + lane);
        if (wide) bias_w[lane + 32] = __ldg(g.bias + n0 + 32 + lane);
        __syncwarp();
      }
      auto issue_aux = [&](int j) {  // box j of the tile this epilogue combines with -> staging
        if (lane == 0) {
          tma_store_wait_read();
          fence_proxy_async_all();
          mbar_expect_tx(&aux_bar[we], (EPI == tc::EPI_DGRAD_HIDDEN ? 2 : 1) * kBoxBytes);
          tma_load_2d(stg, &tm.x0, &aux_bar[we], n0 + 32 * j, xrow);
          if (EPI == tc::EPI_DGRAD_HIDDEN) tma_load_2d(stg + kBoxBytes, &tm.x1, &aux_bar[we], n0 + 32 * j, xrow);
        }
      };
      if (kAux) issue_aux(0);
      float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
      for (int kb0 = 0; kb0 < n_kb; kb0 += g.kseg) {
        const uint32_t hb = hseg & 1;
        bar_wait(&hi_full[hb], (hseg >> 1) & 1, 6);
        tc_fence_after();
        const uint32_t taddr = lane_base + 256 + hb * 128 + cbase;
        if (wide) tmem_accumulate<64>(taddr, acc);
        else tmem_accumulate<32>(taddr, acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(hi_empty_c + hb * 8);
        ++hseg;
      }
      if (three) {
        const uint32_t cb = tcnt & 1;
        bar_wait(&cr_full[cb], (tcnt >> 1) & 1, 7);
        tc_fence_after();
        const uint32_t taddr = lane_base + cb * 128 + cbase;
        if (wide) tmem_accumulate<64>(taddr, acc);
        else tmem_accumulate<32>(taddr, acc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cr_empty_c + cb * 8);
      }
      ++tcnt;
      // ---- finalize: this warp's 32 rows x (64 | 32) columns, one 32-column box at a time ----
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j == 0 || wide) {
          if (kAux) {
            if (j == 1) issue_aux(1);
            bar_wait(&aux_bar[we], aux_cnt & 1, 8);
            ++aux_cnt;
          } else {
            if (lane == 0) tma_store_wait_read();  // the previous box has left the staging area
            __syncwarp();
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = acc[32 * j + 16 * hh + i];
            if (EPI == tc::EPI_FWD_HIDDEN || EPI == tc::EPI_FWD_OUT) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_w + 32 * j + 16 * hh + 4 * j4);
                v[4 * j4 + 0] += b4.x;
                v[4 * j4 + 1] += b4.y;
                v[4 * j4 + 2] += b4.z;
                v[4 * j4 + 3] += b4.w;
              }
              if (g.act == HDPO_ACT_ELU) elu_inplace(v);
              else if (g.act != HDPO_ACT_NONE) act_cold_fwd(g.act, v);
            } else if (EPI == tc::EPI_DGRAD_HIDDEN) {
              float hv[16];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
                const float4 x0 = *reinterpret_cast<const float4*>(stg + lane * 128 + chunk);
                const float4 x1 = *reinterpret_cast<const float4*>(stg + kBoxBytes + lane * 128 + chunk);
                hv[4 * j4 + 0] = x0.x + x1.x;
                hv[4 * j4 + 1] = x0.y + x1.y;
                hv[4 * j4 + 2] = x0.z + x1.z;
                hv[4 * j4 + 3] = x0.w + x1.w;
              }
              if (g.act == HDPO_ACT_ELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= elu_grad_from_out(hv[i]);
              } else if (g.act != HDPO_ACT_NONE) {
                act_cold_bwd(g.act, v, hv);
              }
              if (g.colsum) {
                // bias-gradient partials: column sums over this warp's 32 rows (reduce-scatter butterfly)
                float w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = v[i];
#pragma unroll
                for (int hw = 8, o = 16; hw >= 1; hw >>= 1, o >>= 1) {
                  const bool up = (lane & o) != 0;
#pragma unroll
                  for (int i = 0; i < hw; ++i) {
                    const float send = up ? w[i] : w[i + hw];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, o);
                    w[i] = (up ? w[i + hw] : w[i]) + recv;
                  }
                }
                w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
                if ((lane & 1) == 0) {
                  const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                  g.colsum[(static_cast<size_t>(grow) >> 5) * g.ldc + n0 + 32 * j + 16 * hh + col] = w[0];
                }
              }
            } else if (EPI == tc::EPI_DGRAD_ACCUM) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
                const float4 x0 = *reinterpret_cast<const float4*>(stg + lane * 128 + chunk);
                v[4 * j4 + 0] += x0.x;
                v[4 * j4 + 1] += x0.y;
                v[4 * j4 + 2] += x0.z;
                v[4 * j4 + 3] += x0.w;
              }
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int chunk = ((4 * hh + j4) ^ (lane & 7)) << 4;
              if (kSplit) {
                float4 hi, lo;
                hi.x = tf32_hi(v[4 * j4 + 0]);
                hi.y = tf32_hi(v[4 * j4 + 1]);
                hi.z = tf32_hi(v[4 * j4 + 2]);
                hi.w = tf32_hi(v[4 * j4 + 3]);
                lo.x = tf32_hi(v[4 * j4 + 0] - hi.x);
                lo.y = tf32_hi(v[4 * j4 + 1] - hi.y);
                lo.z = tf32_hi(v[4 * j4 + 2] - hi.z);
                lo.w = tf32_hi(v[4 * j4 + 3] - hi.w);
                *reinterpret_cast<float4*>(stg + lane * 128 + chunk) = hi;
                *reinterpret_cast<float4*>(stg + kBoxBytes + lane * 128 + chunk) = lo;
              } else {
                *reinterpret_cast<float4*>(stg + lane * 128 + chunk) =
                    make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm.c0, stg, n0 + 32 * j, grow);
            if (kSplit) tma_store_2d(&tm.c1, stg + kBoxBytes, n0 + 32 * j, grow);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_read();  // shared memory is released right after the final barrier
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may release shared memory / TMEM the pair's MMAs and commits still use
  if (warp == 1) tmem_dealloc_pair(tmem, kTmemCols);
}

template <int EPI>
static int launch_multi(const tc::GemmTcMaps& tm, const MultiArgs& g, void* stream) {
  auto k = multi_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMultiSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * g.n_pairs);
  cfg.blockDim = dim3(kMultiThreads);
  cfg.dynamicSmemBytes = kMultiSmem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: set-up overlaps the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  HDPO_CUDA_OK(cudaLaunchKernelEx(&cfg, k, tm, g));
  count_launch();
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

static int g_multi_min = -1;
int multi_min_tiles() {
  if (g_multi_min < 0) {
    const char* e = getenv("HDPO_TC_MULTI");
    const char* m = getenv("HDPO_TC_MULTI_MIN");
    // default OFF (measured on B200, 8192 x 512 x 512 3xTF32: 28.2 us vs 33.6 us for the single-tile pairs alone, but
    // the step with four 2048-row chunks on concurrent streams stays ahead: 16.8 ms vs 17.9 ms)
    g_multi_min = (e && atoi(e) != 0) ? (m ? atoi(m) : 48) : (1 << 30);
  }
  return g_multi_min;
}
void set_multi_min_tiles(int min_tiles) { g_multi_min = min_tiles == 0 ? (1 << 30) : min_tiles; }

int gemm_multi(const tc::GemmTcMaps& tm, const tc::GemmTcArgs& a, int epi, int bn, void* stream) {
  HDPO_REQUIRE(bn == 128 || bn == 64, "multi-tile GEMM: tile width %d", bn);
  HDPO_REQUIRE(a.M % kRowTile == 0 && a.N % bn == 0 && a.K % kBK == 0 && a.K > 0 && a.M > 0,
               "multi-tile GEMM shape %dx%dx%d not tileable", a.M, a.N, a.K);
  HDPO_REQUIRE(a.n_pass == 1 || a.n_pass == 3, "n_pass must be 1 or 3");
  static int kseg = 0, max_pairs = 0;
  if (!kseg) {
    kseg = env_int("HDPO_WP_KSEG", 8);
    if (kseg < 1) kseg = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = env_int("HDPO_MULTI_PAIRS", sms / 2);
    if (max_pairs < 1) max_pairs = 1;
  }
  MultiArgs g{};
  g.M = a.M;
  g.N = a.N;
  g.K = a.K;
  g.bn = bn;
  g.n_pass = a.n_pass;
  g.kseg = kseg;
  const int n_tiles = (a.M / kRowTile) * (a.N / bn);
  const int waves = (n_tiles + max_pairs - 1) / max_pairs;
  g.n_pairs = (n_tiles + waves - 1) / waves;  // every pair walks `waves` tiles (the last ones one fewer)
  g.a_row0 = a.a_row0;
  g.b_row0 = a.b_row0;
  g.c_row0 = a.c_row0;
  g.x_row0 = a.x_row0;
  g.ldc = a.ldc;
  g.act = a.act;
  g.bias = a.bias;
  g.colsum = a.colsum_part;
  switch (epi) {
    case tc::EPI_FWD_HIDDEN: return launch_multi<tc::EPI_FWD_HIDDEN>(tm, g, stream);
    case tc::EPI_FWD_OUT: return launch_multi<tc::EPI_FWD_OUT>(tm, g, stream);
    case tc::EPI_DGRAD_HIDDEN: return launch_multi<tc::EPI_DGRAD_HIDDEN>(tm, g, stream);
    case tc::EPI_DGRAD_ACCUM: return launch_multi<tc::EPI_DGRAD_ACCUM>(tm, g, stream);
    case tc::EPI_STORE: return launch_multi<tc::EPI_STORE>(tm, g, stream);
  }
  set_error("bad epilogue %d", epi);
  return HDPO_E_INVALID;
}

// ------------------------------------------------------------------------------------------------------------
// device-side list builder: thread = pair
// ------------------------------------------------------------------------------------------------------------
__global__ void build_tasks_kernel(Sched s, Task* tasks, Task* head_tasks, int n_pairs, int n_servers) {
  pdl_wait();
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_pairs + n_servers) return;
  const int total = s.T * s.nsteps;
  const int u_end = total + (s.k - 1) * s.skew;
  Task end;
  end.a = make_int4(0, 0, 0, 0);
  end.b = make_int4(0, 0, 0, 0);
  if (id >= n_pairs) {
    // ---- list of a head server (8 warps of a policy-head CTA pair): the head tasks of its row tiles, in time order
    const int hp = id - n_pairs;
    Task* hout = head_tasks + static_cast<size_t>(hp) * s.max_tasks;
    int nh = 0;
    for (int u = 0; u < u_end; ++u) {
      for (int c = 0; c < s.k; ++c) {
        const int v = u - c * s.skew;
        if (v < 0 || v >= total) continue;
        const int tt = v / s.nsteps, i = v % s.nsteps;
        if (i != (s.bwd ? 0 : s.nsteps - 1)) continue;
        const int t = s.bwd ? s.T - 1 - tt : tt;
        for (int gi = 0; gi < s.groups; ++gi) {
          const int rt = gi * s.k + c;
          if (rt >= s.R || rt % n_servers != hp) continue;
          Task h;
          if (s.bwd) {
            // adjoint: the head opens the period (it needs the state adjoint completed by the previous sweep position)
            h.a = make_int4(1, t, rt, tt == 0 ? -1 : flag_index(s, tt - 1, s.nsteps - 1, rt));
          } else {
            h.a = make_int4(1, t, rt, flag_index(s, tt, s.nsteps - 1, rt));
          }
          h.b = make_int4(16 * s.C[s.nsteps - 1], flag_index(s, tt, s.nsteps, rt), 0, 0);
          hout[nh++] = h;
        }
      }
    }
    hout[nh] = end;
    return;
  }
  const int pair = id;
  const int gi = pair / s.g, q = pair % s.g;
  Task* out = tasks + static_cast<size_t>(pair) * s.max_tasks;
  int n = 0;
  const int head_count = kServerWarps;  // every warp of the head server arrives once
  for (int u = 0; u < u_end && gi < s.groups; ++u) {
    for (int c = 0; c < s.k; ++c) {
      const int v = u - c * s.skew;
      const int rt = gi * s.k + c;
      if (v < 0 || v >= total || rt >= s.R) continue;
      const int tt = v / s.nsteps, i = v % s.nsteps;
      const int t = s.bwd ? s.T - 1 - tt : tt;
      for (int cc = 0; cc < s.C[i]; ++cc) {
        if (tile_owner(s, c, i, cc, tt) != q) continue;
        Task k;
        k.a = make_int4(1 | (s.layer[i] << 8) | (s.epi[i] << 16), t, rt, cc * s.bn[i]);
        int wf, wc;
        if (i > 0) {
          wf = flag_index(s, tt, i - 1, rt);
          wc = 16 * s.C[i - 1];
        } else if (s.bwd) {
          wf = flag_index(s, tt, s.nsteps, rt);
          wc = head_count;
        } else {
          wf = tt == 0 ? -1 : flag_index(s, tt - 1, s.nsteps, rt);
          wc = head_count;
        }
        k.b = make_int4(s.nkb[i], wf, wc, flag_index(s, tt, i, rt));
        out[n++] = k;
      }
    }
  }
  out[n] = end;
}

// ------------------------------------------------------------------------------------------------------------
// pre-pass kernels: transposed copies the per-thread head reads coalesced
// ------------------------------------------------------------------------------------------------------------
// dst[c][b] = src[b][c] for b < rows_valid (else 0); src row-major with leading dimension ld_src, c < cols
__global__ void __launch_bounds__(256) transpose_rows_kernel(const float* __restrict__ src, int rows_valid, int Bp, int cols,
                                                             int ld_src, float* __restrict__ dst) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int b0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r, c = c0 + tx;
    tile[r][tx] = (src && b < rows_valid && c < cols) ? src[static_cast<size_t>(b) * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, b = b0 + tx;
    if (c < cols && b < Bp) dst[static_cast<size_t>(c) * Bp + b] = tile[tx][r];
  }
}

// dT[t][s][b] = demand of (b, s) in period t (column t + period_shift of the input), 0 for padding rows
__global__ void __launch_bounds__(256) demand_t_kernel(const float* __restrict__ demands, int layout, int B, int Bp, int S, int T,
                                                       int t_stride, int shift, int B_total, float* __restrict__ dT) {
  pdl_wait();
  if (layout == HDPO_DEMAND_TSB) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(T) * S * Bp) return;
    const int b = static_cast<int>(i % Bp);
    const size_t ts = i / Bp;
    const int s = static_cast<int>(ts % S), t = static_cast<int>(ts / S);
    dT[i] = b < B ? demands[(static_cast<size_t>(t + shift) * S + s) * B_total + b] : 0.f;
    return;
  }
  // BST: block = (32 scenarios) x (32 periods) of one store, transposed through shared memory
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nbt = (T + 31) / 32;
  const int bt = blockIdx.x % nbt;
  const int rest = blockIdx.x / nbt;
  const int s = rest % S, bb = rest / S;
  const int b0 = bb * 32, t0 = bt * 32;
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r, t = t0 + tx;
    tile[r][tx] = (b < B && t < T) ? demands[(static_cast<size_t>(b) * S + s) * t_stride + t + shift] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, b = b0 + tx;
    if (t < T && b < Bp) dT[(static_cast<size_t>(t) * S + s) * Bp + b] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256) zero_ints_kernel(int* __restrict__ p, size_t n) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}
__global__ void __launch_bounds__(256) zero_floats_kernel(float* __restrict__ p, size_t n) {
  pdl_wait();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static unsigned long long* g_trace = nullptr;
static int g_trace_cap = 0;
void set_trace(unsigned long long* buf, int cap_per_role) {
  g_trace = buf;
  g_trace_cap = cap_per_role;
}

// Opt-in (HDPO_WIDE_PERSIST=1 or hdpo_debug_set_wide_persist): measured on B200 the one-launch sweeps are parity-green
// but SLOWER than the per-period chain of rollout_wide.cu at 8192 scenarios (see DESIGN.md "persistent sweeps").
static int g_enabled = -1;
bool enabled() {
  if (g_enabled < 0) g_enabled = env_int("HDPO_WIDE_PERSIST", 0) != 0;
  return g_enabled != 0;
}
void set_enabled(int on) { g_enabled = on != 0; }
bool eligible(const HdpoRolloutDesc* d) {
  if (d->arch != HDPO_ARCH_VANILLA_WAREHOUSE || d->precision == HDPO_PREC_FP32) return false;
  if (d->pb.W < 1 || d->pb.W > kMaxW || d->pb.E != 0) return false;
  if (d->master.n_layers < 2 || d->master.n_layers > HDPO_MAX_LAYERS) return false;
  if (d->T > 30000 ||  d->pb.L < 2 || d->pb.Lw < 2) return false;
  const int adj_floats = d->pb.W > 1 ? ((d->pb.S * d->pb.W + 3) & ~3) : 0;
  const int budget = (kHeadBytes / 4 - adj_floats) / kHeadWarps / 4 * 4;  // floats of shared memory per head warp
  if (head_warp_floats<false>(d->pb.S, d->pb.W, d->pb.L, d->pb.Lw) > budget) return false;
  if (head_warp_floats<true>(d->pb.S, d->pb.W, d->pb.L, d->pb.Lw) > budget) return false;
  return true;
}

static size_t a256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static int tile_bn(int width) { return width % 128 == 0 ? 128 : 64; }

// groups of `g` pairs, `k` row tiles (chains) per group
static void choose_groups(Sched* s, int cmax) {
  const int max_pairs = env_int("HDPO_WP_PAIRS", kMaxPairs) < kMaxPairs ? env_int("HDPO_WP_PAIRS", kMaxPairs) : kMaxPairs;
  int g = cmax;
  if (g > max_pairs) g = max_pairs;
  if (g < 1) g = 1;
  int groups = max_pairs / g;
  if (groups > s->R) groups = s->R;
  int k = (s->R + groups - 1) / groups;
  groups = (s->R + k - 1) / k;
  s->g = g;
  s->k = k;
  s->groups = groups;
}

// tile shapes of one direction: forward step l = layer l ([rows x wp[l]] x W_l^T -> wp[l+1] columns); adjoint step i =
// dgrad of layer l = n-1-i ([rows x wp[l+1]] x WT_l^T -> wp[l] columns). Returns the list stride a pair needs.
static void fill_sched(Sched* s, int T, int Bp, const int* wp, int n, int bwd) {
  s->T = T;
  s->nsteps = n;
  s->R = Bp / kRowTile;
  s->bwd = bwd;
  int cmax = 1, csum = 0;
  for (int i = 0; i < n; ++i) {
    const int l = bwd ? n - 1 - i : i;
    const int width = bwd ? wp[l] : wp[l + 1];
    const bool narrow = bwd ? l == 0 : l + 1 == n;  // output layer / state adjoint: 64-column tiles
    const int bn = narrow ? 64 : tile_bn(width);
    s->C[i] = width / bn;
    s->bn[i] = bn;
    s->nkb[i] = (bwd ? wp[l + 1] : wp[l]) / kBK;
    s->layer[i] = i;
    s->epi[i] = bwd ? (l > 0 ? EPI_DGRAD_HIDDEN : EPI_DGRAD_GX) : (l + 1 < n ? EPI_FWD_HIDDEN : EPI_FWD_OUT);
    if (s->C[i] > cmax) cmax = s->C[i];
    csum += s->C[i];
  }
  choose_groups(s, cmax);
  s->skew = env_int("HDPO_WP_SKEW", n / 2);
  if (s->k < 2) s->skew = 0;
  s->max_tasks = T * s->k * csum + 8;  // even if one pair of a group owned every tile of its chains
  const int min_servers = kMinHeadPairs * kServersPerPair;
  const int head_list = T * ((s->R + min_servers - 1) / min_servers) + 8;  // list of a head server
  if (head_list > s->max_tasks) s->max_tasks = head_list;
}

struct Extra {
  size_t o_XT, o_OUTT, o_hT, o_pT, o_ltT, o_whT, o_dT, o_gXT, o_gYT, o_flags, o_tasks, o_htasks, total;
  size_t n_flags;
  int max_tasks;
};
static Extra plan_extra(int T, int S, int W, int Bp, const int* wp, int n, int save) {
  Extra e;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += a256(bytes);
    return at;
  };
  const size_t f = sizeof(float), B = static_cast<size_t>(Bp);
  const int ldx = wp[0], ldy = wp[n];
  e.o_XT = take(static_cast<size_t>(save ? T + 1 : 2) * ldx * B * f);
  e.o_OUTT = take(static_cast<size_t>(save ? T : 1) * ldy * B * f);
  e.o_hT = take(static_cast<size_t>(S) * B * f);
  e.o_pT = take(static_cast<size_t>(S) * B * f);
  e.o_ltT = take(static_cast<size_t>(S) * W * B * f);
  e.o_whT = take(static_cast<size_t>(3) * W * B * f);
  e.o_dT = take(static_cast<size_t>(T) * S * B * f);
  e.o_gXT = take(save ? static_cast<size_t>(ldx) * B * f : 0);
  e.o_gYT = take(save ? static_cast<size_t>(ldy) * B * f : 0);
  const int R = Bp / kRowTile;
  e.n_flags = static_cast<size_t>(T) * (n + 1) * R;
  e.o_flags = take(e.n_flags * sizeof(int));
  Sched sf{}, sb{};
  fill_sched(&sf, T, Bp, wp, n, 0);
  fill_sched(&sb, T, Bp, wp, n, 1);
  e.max_tasks = sf.max_tasks > sb.max_tasks ? sf.max_tasks : sb.max_tasks;
  e.o_tasks = take(static_cast<size_t>(kMaxPairs) * e.max_tasks * sizeof(Task));
  e.o_htasks = take(static_cast<size_t>(74 * kServersPerPair) * e.max_tasks * sizeof(Task));
  e.total = o + 256;
  return e;
}

size_t extra_bytes(const HdpoRolloutDesc* d, int Bp, const int* wp, int n) {
  return plan_extra(d->T, d->pb.S, d->pb.W, Bp, wp, n, d->save_for_backward).total;
}

static float* xf(const Ctx& c, size_t off) { return reinterpret_cast<float*>(static_cast<char*>(c.extra) + off); }

static HeadP make_head(const Ctx& c, const Extra& e) {
  HeadP h{};
  h.B = c.B;
  h.Bp = c.Bp;
  h.S = c.S;
  h.W = c.W;
  h.L = c.L;
  h.Lw = c.Lw;
  h.ldx = c.wp[0];
  h.ldy = c.wp[c.n];
  h.nS = c.S * c.L;
  h.T = c.T;
  h.lost = c.lost;
  h.profit = c.profit;
  h.has_edge = c.has_edge;
  h.transshipment = c.transshipment;
  h.discrete = c.discrete;
  h.save = c.save;
  h.ignore_periods = c.ignore_periods;
  h.B_total = c.B_total;
  h.wub = c.wub;
  h.adjacency = c.W > 1 ? c.adjacency : nullptr;
  h.dT = xf(c, e.o_dT);
  h.hT = xf(c, e.o_hT);
  h.pT = xf(c, e.o_pT);
  h.ltT = xf(c, e.o_ltT);
  h.whT = xf(c, e.o_whT);
  h.XT = xf(c, e.o_XT);
  h.OUTT = xf(c, e.o_OUTT);
  h.X = c.X;
  h.X_hi = c.X_hi;
  h.X_lo = c.X_lo;
  h.gXT = xf(c, e.o_gXT);
  h.gYT = xf(c, e.o_gYT);
  h.gY_hi = c.gz_hi[c.n - 1];
  h.gY_lo = c.gz_lo[c.n - 1];
  h.cost_b = c.cost_b;
  h.report_b = c.report_b;
  h.reward_tb = c.reward_tb;
  return h;
}

static int head_pairs_for(int n_pairs, int* out) {
  int dev = 0, sms = 0;
  HDPO_CUDA_OK(cudaGetDevice(&dev));
  HDPO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int hp = (sms - 2 * n_pairs) / 2;  // one CTA per SM, all co-resident: the dependency flags need every CTA running
  HDPO_REQUIRE(hp >= kMinHeadPairs, "persistent wide path: %d SMs leave no room for the policy-head CTAs next to %d GEMM pairs", sms,
               n_pairs);
  *out = hp;
  return HDPO_OK;
}

template <bool BWD>
static int launch_persist(const Params& p, int n_pairs, int n_hpairs, void* stream) {
  auto k = persist_kernel<BWD>;
  static bool configured = false;
  if (!configured) {
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (n_pairs + n_hpairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemTotal;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  HDPO_CUDA_OK(cudaLaunchKernelEx(&cfg, k, p));
  count_launch();
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

static int make_pair_maps(CUtensorMap* hi, CUtensorMap* lo, const float* phi, const float* plo, uint64_t rows, uint64_t cols,
                          uint32_t box_rows) {
  int rc = make_tensor_map(hi, phi, rows, cols, cols, box_rows);
  if (rc) return rc;
  return make_tensor_map(lo, plo ? plo : phi, rows, cols, cols, box_rows);
}

static int prepass(const Ctx& c, const Extra& e) {
  void* stream = c.stream;
  const int Bp = c.Bp;
  auto tr = transpose_rows_kernel;
  auto launch_t = [&](const float* src, int cols, int ld, float* dst) {
    HDPO_LAUNCH_PDL(tr, dim3(Bp / 32, (cols + 31) / 32), 256, 0, stream, src, c.B, Bp, cols, ld, dst);
  };
  launch_t(c.st.holding_costs, c.S, c.S, xf(c, e.o_hT));
  launch_t(c.st.underage_costs, c.S, c.S, xf(c, e.o_pT));
  launch_t(c.st.lead_times, c.S * c.W, c.S * c.W, xf(c, e.o_ltT));
  launch_t(c.st.warehouse_holding_costs, c.W, c.W, xf(c, e.o_whT));
  launch_t(c.st.warehouse_lead_times, c.W, c.W, xf(c, e.o_whT) + static_cast<size_t>(c.W) * Bp);
  launch_t(c.st.warehouse_edge_costs, c.W, c.W, xf(c, e.o_whT) + static_cast<size_t>(2) * c.W * Bp);
  HDPO_LAUNCH_OK();
  auto dk = demand_t_kernel;
  unsigned blocks;
  if (c.demand_layout == HDPO_DEMAND_TSB) blocks = static_cast<unsigned>(ceil_div64(static_cast<int64_t>(c.T) * c.S * Bp, 256));
  else blocks = static_cast<unsigned>(static_cast<size_t>((c.T + 31) / 32) * c.S * (Bp / 32));
  HDPO_LAUNCH_PDL(dk, blocks, 256, 0, stream, c.demands, c.demand_layout, c.B, Bp, c.S, c.T, c.t_stride, c.period_shift,
                  c.B_total, xf(c, e.o_dT));
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

static int zero_flags(const Ctx& c, const Extra& e) {
  auto zk = zero_ints_kernel;
  HDPO_LAUNCH_PDL(zk, static_cast<unsigned>(ceil_div64(static_cast<int64_t>(e.n_flags), 256)), 256, 0, c.stream,
                  reinterpret_cast<int*>(static_cast<char*>(c.extra) + e.o_flags), e.n_flags);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

static int build_lists(const Ctx& c, const Extra& e, const Sched& s, int n_pairs, int n_hpairs) {
  auto bk = build_tasks_kernel;
  const int n_servers = n_hpairs * kServersPerPair;
  HDPO_LAUNCH_PDL(bk, (n_pairs + n_servers + 31) / 32, 32, 0, c.stream, s,
                  reinterpret_cast<Task*>(static_cast<char*>(c.extra) + e.o_tasks),
                  reinterpret_cast<Task*>(static_cast<char*>(c.extra) + e.o_htasks), n_pairs, n_servers);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

int forward(const Ctx& c) {
  const Extra e = plan_extra(c.T, c.S, c.W, c.Bp, c.wp, c.n, c.save);
  HDPO_REQUIRE(c.Bp % kRowTile == 0, "persistent wide path: rows must be padded to %d", kRowTile);
  Params p{};
  Sched s{};
  fill_sched(&s, c.T, c.Bp, c.wp, c.n, 0);
  s.max_tasks = e.max_tasks;
  const uint64_t tslots = c.save ? static_cast<uint64_t>(c.T) : 1;
  for (int l = 0; l < c.n; ++l) {
    const bool hidden = l + 1 < c.n;
    const int bn = s.bn[l];
    LayerDesc& L = p.L[l];
    const float* a_hi = l == 0 ? c.X_hi : c.act_hi[l - 1];
    const float* a_lo = l == 0 ? c.X_lo : c.act_lo[l - 1];
    int rc = make_pair_maps(&L.a_hi, &L.a_lo, a_hi, a_lo, tslots * c.Bp, c.wp[l], 128);
    if (!rc) rc = make_pair_maps(&L.b_hi, &L.b_lo, c.W_hi[l], c.W_lo[l], c.wp[l + 1], c.wp[l], bn / 2);
    if (!rc) rc = make_pair_maps(&L.c_hi, &L.c_lo, c.act_hi[l], hidden ? c.act_lo[l] : nullptr, tslots * c.Bp, c.wp[l + 1], 32);
    if (rc) return rc;
    L.x_hi = L.c_hi;
    L.x_lo = L.c_lo;
    L.bias = c.bias[l];
    L.colsum = nullptr;
    L.act = c.act[l];
    L.bn = bn;
    L.ldc = c.wp[l + 1];
    L.a_tmul = L.c_tmul = L.x_tmul = c.save ? c.Bp : 0;
  }
  const int n_pairs = s.groups * s.g;
  HDPO_REQUIRE(n_pairs <= kMaxPairs, "persistent wide path: %d pairs do not fit", n_pairs);
  p.head = make_head(c, e);
  p.tasks = reinterpret_cast<const Task*>(static_cast<char*>(c.extra) + e.o_tasks);
  p.head_tasks = reinterpret_cast<const Task*>(static_cast<char*>(c.extra) + e.o_htasks);
  p.flags = reinterpret_cast<int*>(static_cast<char*>(c.extra) + e.o_flags);
  p.max_tasks = e.max_tasks;
  p.n_pass = c.n_pass;
  p.kseg = env_int("HDPO_WP_KSEG", 8);
  if (p.kseg < 1) p.kseg = 1;
  p.n_pairs = n_pairs;
  p.trace = g_trace;
  p.trace_cap = g_trace_cap;
  int rc;
  if ((rc = prepass(c, e))) return rc;
  {
    // transposed copy of the initial state (X[0] row-major was written by init_state_kernel)
    auto tr = transpose_rows_kernel;
    HDPO_LAUNCH_PDL(tr, dim3(c.Bp / 32, (c.wp[0] + 31) / 32), 256, 0, c.stream, static_cast<const float*>(c.X), c.Bp, c.Bp,
                    c.wp[0], c.wp[0], xf(c, e.o_XT));
    HDPO_LAUNCH_OK();
  }
  if ((rc = zero_flags(c, e))) return rc;
  int n_hpairs = 0;
  if ((rc = head_pairs_for(n_pairs, &n_hpairs))) return rc;
  if ((rc = build_lists(c, e, s, n_pairs, n_hpairs))) return rc;
  return launch_persist<false>(p, n_pairs, n_hpairs, c.stream);
}

int backward(const Ctx& c, float g_total, float g_report) {
  const Extra e = plan_extra(c.T, c.S, c.W, c.Bp, c.wp, c.n, c.save);
  HDPO_REQUIRE(c.save, "the adjoint sweep needs the tapes of a forward run with save_for_backward");
  Params p{};
  Sched s{};
  fill_sched(&s, c.T, c.Bp, c.wp, c.n, 1);
  s.max_tasks = e.max_tasks;
  const uint64_t rows = static_cast<uint64_t>(c.T) * c.Bp;
  for (int i = 0; i < c.n; ++i) {
    const int l = c.n - 1 - i;  // step i = dgrad of layer l: [rows x wp[l+1]] x WT_l [wp[l] x wp[l+1]]^T -> [rows x wp[l]]
    const int bn = s.bn[i];
    LayerDesc& L = p.L[i];
    int rc = make_pair_maps(&L.a_hi, &L.a_lo, c.gz_hi[l], c.gz_lo[l], rows, c.wp[l + 1], 128);
    if (!rc) rc = make_pair_maps(&L.b_hi, &L.b_lo, c.WT_hi[l], c.WT_lo[l], c.wp[l], c.wp[l + 1], bn / 2);
    if (!rc && l > 0) {
      rc = make_pair_maps(&L.c_hi, &L.c_lo, c.gz_hi[l - 1], c.gz_lo[l - 1], rows, c.wp[l], 32);
      if (!rc) rc = make_pair_maps(&L.x_hi, &L.x_lo, c.act_hi[l - 1], c.act_lo[l - 1], rows, c.wp[l], 32);
    }
    if (rc) return rc;
    if (l == 0) L.c_hi = L.c_lo = L.x_hi = L.x_lo = L.a_hi;  // unused
    L.bias = nullptr;
    L.colsum = l > 0 ? c.csum[l - 1] : nullptr;
    L.act = l > 0 ? c.act[l - 1] : HDPO_ACT_NONE;
    L.bn = bn;
    L.ldc = c.wp[l];
    L.a_tmul = c.Bp;
    L.c_tmul = L.x_tmul = l > 0 ? c.Bp : 0;
  }
  const int n_pairs = s.groups * s.g;
  HDPO_REQUIRE(n_pairs <= kMaxPairs, "persistent wide path: %d pairs do not fit", n_pairs);
  p.head = make_head(c, e);
  p.head.g_total = g_total;
  p.head.g_report = g_report;
  p.tasks = reinterpret_cast<const Task*>(static_cast<char*>(c.extra) + e.o_tasks);
  p.head_tasks = reinterpret_cast<const Task*>(static_cast<char*>(c.extra) + e.o_htasks);
  p.flags = reinterpret_cast<int*>(static_cast<char*>(c.extra) + e.o_flags);
  p.max_tasks = e.max_tasks;
  p.n_pass = c.n_pass;
  p.kseg = env_int("HDPO_WP_KSEG", 8);
  if (p.kseg < 1) p.kseg = 1;
  p.n_pairs = n_pairs;
  p.trace = g_trace;
  p.trace_cap = g_trace_cap;
  int rc;
  {
    auto zk = zero_floats_kernel;
    const size_t n = static_cast<size_t>(c.wp[0]) * c.Bp;
    HDPO_LAUNCH_PDL(zk, static_cast<unsigned>(ceil_div64(static_cast<int64_t>(n), 256)), 256, 0, c.stream, xf(c, e.o_gXT), n);
    HDPO_LAUNCH_OK();
  }
  if ((rc = zero_flags(c, e))) return rc;
  int n_hpairs = 0;
  if ((rc = head_pairs_for(n_pairs, &n_hpairs))) return rc;
  if ((rc = build_lists(c, e, s, n_pairs, n_hpairs))) return rc;
  return launch_persist<true>(p, n_pairs, n_hpairs, c.stream);
}

}  // namespace wp
}  // namespace hdpo
#endif
