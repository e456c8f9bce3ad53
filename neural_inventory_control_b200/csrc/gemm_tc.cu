// gemm_tc.cu - tcgen05 / TMEM / TMA tile GEMM (see gemm_tc.cuh). Inline PTX only; no CUTLASS dependency.
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"
#include "wide_persist.cuh"

#include <cstdlib>

#ifndef HDPO_EMU

// A/B build switches (tools/epi_ab.sh): packed f32x2 accumulator sums / bias / split in the epilogue, packed ELU.
// Measured on B200 (8192 x 50 x 50 stores, 3xTF32, three interleaved runs each): all scalar 16.18 - 16.20 ms per step, packed
// ELU only 16.23 - 16.24, packed sums only 16.31 - 16.32, both 16.35 - 16.37: a quarter fewer issued instructions and a small
// LOSS (the epilogue warps wait on TMEM loads, MUFU and the store drain; the pack / unpack moves lengthen the chains).
// The scalar forms are the default here; the SIMT kernels (small nets, SymmetryAware heads) keep the packed ELU (+1 %).
#ifndef HDPO_EPI_PACK_SUM
#define HDPO_EPI_PACK_SUM 0
#endif
#ifndef HDPO_EPI_PACK_ELU
#define HDPO_EPI_PACK_ELU 0
#endif
#ifndef HDPO_EPI_PIPE
#define HDPO_EPI_PIPE 0
#endif
// where a CTA lets the dependent kernel of its stream start (griddepcontrol.launch_dependents): 0 = right after its own
// dependency wait, 1 = once its accumulators are complete, 2 = after its epilogue, 3 = never (implicit at grid completion),
// -1 (default) = 0 for the forward epilogues and GemmTcArgs::pdl_late for the adjoint ones (the wide rollout sets it when
// several chunk streams compete for the SMs). Measured on B200 (8192 x 50 x 50 stores = 4 chunks, three interleaved runs, ms
// per step / forward / adjoint): 0: 16.16 / 5.36 / 10.81; 1: 16.06 - 16.09 / 5.40 / 10.69; 2: 16.09 - 16.15 / 5.39 / 10.72 - 10.80;
// 3: 16.19 - 16.24 / 5.43 / 10.76 (a dependent CTA that was launched early holds an SM - 198 KB of shared memory, all of
// TMEM - while it waits, which the other chunk streams could have used). With ONE chunk (latency-bound chain, idle SMs)
// the early trigger is better: many_warehouses 1024 scenarios adjoint 3.82 (early) vs 4.05 ms (late).
#ifndef HDPO_PDL_TRIGGER
#define HDPO_PDL_TRIGGER -1
#endif

namespace hdpo {
namespace tc {

// ------------------------------------------------------------------------------------------------------------
// kernel: 10 warps = TMA producer | MMA issuer (+TMEM owner) | 8 epilogue warps (128 rows x 2 column halves)
// ------------------------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 8;                  // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA producer warp + MMA warp + epilogue warps
constexpr int kBK = 32;  // floats per K block = one 128-byte swizzle row
// Accumulator splitting. The tensor core adds each K=8 product block into the fp32 accumulator with truncation, so
// the error of ONE accumulator grows linearly with the number of tcgen05.mma issued into it (measured: 4e-6
// relative at K = 512, 192 MMAs). We therefore keep 4 accumulators in TMEM: [0] collects the small cross terms
// (lo*hi, hi*lo; 2^-11 of the magnitude, their truncation is negligible) and [1..3] each take one third of the K
// range of the hi*hi term; the epilogue adds the four in fp32 with round-to-nearest. 4 x 128 columns = all of TMEM.
constexpr int kAccums = 4;
constexpr int kHiChunks = kAccums - 1;

// CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BN tile; each CTA stages its own 128 A rows
// and HALF of the B rows, so the shared-memory traffic (TMA fills + MMA operand reads, the measured mainloop bound of
// the single-CTA form) and the L2 traffic per flop drop by a quarter.
// EARLY_AUX (CTA-pair dgrad epilogue): one stage less, and 64 KB after the ring receive half of the saved-activation
// tile before the mainloop starts (the other half is fetched into the idle ring once the accumulators are ready).
// OCC = 2: two CTAs per SM (BN = 64 only: 4 x 64 = 256 TMEM columns and a 2-stage ring each), so that the ramp, epilogue
// and teardown of one tile overlap the MMAs of the tile that shares the SM - the hardware scheduler does the
// interleaving across the concurrent chunk streams that a hand-written persistent loop could not.
template <int BN, int CG = 1, bool EARLY_AUX = false, int OCC = 1>
struct SmemPlan {
  static constexpr int kStages = OCC == 2 ? 2 : ((BN == 128 && CG == 1) ? 3 : (EARLY_AUX ? 3 : 4));
  static constexpr int kAuxBytes = EARLY_AUX ? kEpiWarps * 2 * 32 * 128 : 0;
  static constexpr int kABytes = 128 * kBK * 4;
  static constexpr int kBBytes = (BN / CG) * kBK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kRing = kStages * kStageBytes;
  static constexpr int kTotal = kRing + kAuxBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + BN * 4 /*bias tile*/;
};

// MN = false: D[M,N] = A[M,K] B[N,K]^T, operands K-major (rows = m / n, contiguous k).
// MN = true : D[M,N] = sum_k A[k][m] B[k][n], operands MN-major (rows = k, contiguous m / n): the weight-gradient
//             form dW = gz^T h straight from the [row][feature] tapes; blockIdx.z selects a K range of k_per_split rows
//             and writes its own partial slice (short ranges keep the truncating accumulation fp32-grade).
template <int BN, int EPI, bool MN, int CG = 1, int OCC = 1>
__global__ void __launch_bounds__(kThreads, OCC)
gemm_tc_kernel(const __grid_constant__ GemmTcMaps tm, GemmTcArgs g) {
  // (MN && CG == 2: weight-gradient tiles of 256 x BN per CTA pair - each CTA stages its own 128 m-columns of A and HALF
  // of the n-columns of B: a quarter fewer operand bytes per flop than the single-CTA form, which is L2-bound)
  static_assert(OCC == 1 || (OCC == 2 && BN == 64 && !MN), "two CTAs per SM: 64-column K-major tiles only");
  const CUtensorMap& tm_a_hi = tm.a_hi;
  const CUtensorMap& tm_a_lo = tm.a_lo;
  const CUtensorMap& tm_b_hi = tm.b_hi;
  const CUtensorMap& tm_b_lo = tm.b_lo;
  // Measured on B200 (cfg 4, in the pipeline): with the early half-tile the dgrad epilogue shrinks 4.25 -> 3.6 us but
  // the 3-stage ring lengthens the mainloop 6.6 -> 7.7 us, a net loss; the path is kept for A/B runs only.
  constexpr bool kEarlyAux = false && CG == 2 && EPI == EPI_DGRAD_HIDDEN;
  using P = SmemPlan<BN, CG, kEarlyAux, OCC>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);  // SW128 needs 1024-byte aligned tiles
  unsigned char* aux_early = smem + P::kRing;  // [epilogue warp][hi | lo][32 rows x 128 B], kEarlyAux only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P::kRing + P::kAuxBytes);
  uint64_t* empty_bar = full_bar + P::kStages;
  uint64_t* accum_bar = empty_bar + P::kStages;
  uint64_t* epi_bar = accum_bar + 1;  // one per epilogue warp: its TMA loads of the tile the epilogue combines with
  uint64_t* aux_bar = epi_bar + kEpiWarps;  // one per epilogue warp: the early half of the combined tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + kEpiWarps);
  float* bias_s = reinterpret_cast<float*>(smem + P::kRing + P::kAuxBytes + 256);  // [BN], forward epilogues

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CG = 1: grid (N tiles, M tiles, K slices); CG = 2: grid (2 = CTA of the pair, 256-row tiles, N tiles)
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  // MN && CG == 2: grid (2, row tiles x column tiles, K slices)
  const int pair_tile_n = (MN && CG == 2) ? static_cast<int>(blockIdx.y) % (g.N / BN) : 0;
  const int pair_tile_m = (MN && CG == 2) ? static_cast<int>(blockIdx.y) / (g.N / BN) : static_cast<int>(blockIdx.y);
  const int m0 = CG == 2 ? pair_tile_m * 256 + static_cast<int>(rank) * 128 : blockIdx.y * 128;
  const int n0 = CG == 2 ? (MN ? pair_tile_n : static_cast<int>(blockIdx.z)) * BN : blockIdx.x * BN;
  long long* dbg = g.dbg_clock ? g.dbg_clock + 8 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  // K slices: rows of the MN-major operands (weight gradient) / columns of the K-major ones (split-K, single CTA only)
  const bool ksplit = MN || (CG == 1 && g.k_per_split > 0);
  const int n_kb = (ksplit ? g.k_per_split : g.K) / kBK;
  const int k_begin = ksplit ? static_cast<int>(blockIdx.z) * g.k_per_split : 0;
  const bool three = g.n_pass == 3;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_b_hi);
    if (three) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_b_lo);
    }
    for (int s = 0; s < P::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    for (int w = 0; w < kEpiWarps; ++w) {
      mbar_init(&epi_bar[w], 1);
      mbar_init(&aux_bar[w], 1);
    }
    prefetch_tmap(&tm.c0);
    if (EPI == EPI_FWD_HIDDEN || EPI == EPI_DGRAD_HIDDEN) prefetch_tmap(&tm.c1);
    if (EPI == EPI_DGRAD_HIDDEN || EPI == EPI_DGRAD_ACCUM) prefetch_tmap(&tm.x0);
    if (EPI == EPI_DGRAD_HIDDEN) prefetch_tmap(&tm.x1);
    fence_barrier_init();
  }
  // Inputs that were produced at least two kernels back are touched BEFORE the dependency wait (and before the set-up barrier
  // below, which also publishes the bias tile to the other epilogue warps), so that their latency
  // (several microseconds when the memory system is busy) overlaps the predecessor: the bias tile goes to shared
  // memory, the saved-activation tile of the dgrad epilogue is pulled into L2.
  if (warp >= 2) {
    const int we0 = warp - 2, q0 = warp & 3, half0 = we0 >> 2;
    if (EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_OUT) {
      const int t = threadIdx.x - 64;
      if (t < BN) bias_s[t] = (CG == 1 && !MN && blockIdx.z > 0) ? 0.f : g.bias[n0 + t];  // split-K: slice 0 adds it
    }
    if (EPI == EPI_DGRAD_HIDDEN && lane == 0) {
      const int xrow = g.x_row0 + m0 + q0 * 32;
#pragma unroll
      for (int b = kEarlyAux ? 1 : 0; b < BN / 64; ++b) {
        tma_prefetch_l2_2d(&tm.x0, n0 + half0 * (BN / 2) + 32 * b, xrow);
        tma_prefetch_l2_2d(&tm.x1, n0 + half0 * (BN / 2) + 32 * b, xrow);
      }
    }
  }
  if (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
  // TMEM: kAccums accumulators of BN fp32 columns each (see "accumulator splitting" at the MMA loop)
  if (warp == 1) {
    if (CG == 2) tmem_alloc_pair(tmem_slot, kAccums * BN);
    else tmem_alloc(tmem_slot, kAccums * BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (kEarlyAux && warp >= 2 && lane == 0) {
    // first 32-column box (hi and lo) of this warp's saved-activation sub-tile: a forward tape, independent of the
    // predecessor kernel, so the load is in flight during the dependency wait and the whole mainloop
    const int we0 = warp - 2, q0 = warp & 3, half0 = we0 >> 2;
    unsigned char* dst = aux_early + we0 * (2 * 32 * 128);
    mbar_expect_tx(&aux_bar[we0], 2 * 32 * 128);
    tma_load_2d(dst, &tm.x0, &aux_bar[we0], n0 + half0 * (BN / 2), g.x_row0 + m0 + q0 * 32);
    tma_load_2d(dst + 32 * 128, &tm.x1, &aux_bar[we0], n0 + half0 * (BN / 2), g.x_row0 + m0 + q0 * 32);
  }
  // everything above overlapped the tail of the previous kernel in the stream (PDL); from here on we read its output
  // (-1: the caller decides per launch - g.pdl_late - and only the adjoint epilogues honour it)
  const int kPdlTrigger = HDPO_PDL_TRIGGER >= 0
                              ? HDPO_PDL_TRIGGER
                              : (((EPI == EPI_DGRAD_HIDDEN || EPI == EPI_DGRAD_ACCUM) && g.pdl_late) ? 1 : 0);
  pdl_wait();
  if (kPdlTrigger == 0) pdl_launch_dependents();
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  // trace records start once the predecessor's data is visible (a PDL-launched CTA may have idled above for long)
  const unsigned long long trace_t0 = (g.trace.buf && threadIdx.x == 0) ? trace_now() : 0ull;

  if (warp == 0) {
    // ===== TMA producer =====
    // CTA pair: the whole warp runs the loop and one elected lane issues (warp-uniform control flow keeps the
    // descriptors in uniform registers; measured: the pair's mainloop drops from ~1000 to ~780 cycles per K block,
    // the MMA floor). Single CTA: lane 0 alone runs the loop (measured faster there: 19.1k vs 20.7k cycles per tile).
    if (CG == 2 || lane == 0) {
      const uint32_t tx = (three ? P::kStageBytes : (P::kABytes + P::kBBytes)) * CG;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % P::kStages;
        const uint32_t ph = (kb / P::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st = smem + s * P::kStageBytes;
        if (CG == 1 || elect_one()) {
        if (CG == 2) {
          // both CTAs fill their own stage; all bytes are counted on the LEADER's full barrier (the MMA issuer's)
          if (rank == 0) mbar_expect_tx(&full_bar[s], tx);
          const uint32_t bar = mapa_u32(&full_bar[s], 0);
          if (MN) {
            constexpr int kBox = kBK * 128;  // bytes of one {32 mn, BK k} box
            const int krow = k_begin + kb * kBK;
            const int ncol = n0 + static_cast<int>(rank) * (BN / 2);
            for (int i = 0; i < 128 / 32; ++i) {
              tma_load_2d_pair(st + i * kBox, &tm_a_hi, bar, m0 + 32 * i, g.a_row0 + krow);
              if (three) tma_load_2d_pair(st + P::kABytes + i * kBox, &tm_a_lo, bar, m0 + 32 * i, g.a_row0 + krow);
            }
            for (int i = 0; i < BN / 2 / 32; ++i) {
              tma_load_2d_pair(st + 2 * P::kABytes + i * kBox, &tm_b_hi, bar, ncol + 32 * i, g.b_row0 + krow);
              if (three)
                tma_load_2d_pair(st + 2 * P::kABytes + P::kBBytes + i * kBox, &tm_b_lo, bar, ncol + 32 * i,
                                 g.b_row0 + krow);
            }
          } else {
          const int brow = g.b_row0 + n0 + static_cast<int>(rank) * (BN / 2);
          tma_load_2d_pair(st, &tm_a_hi, bar, kb * kBK, g.a_row0 + m0);
          tma_load_2d_pair(st + 2 * P::kABytes, &tm_b_hi, bar, kb * kBK, brow);
          if (three) {
            tma_load_2d_pair(st + P::kABytes, &tm_a_lo, bar, kb * kBK, g.a_row0 + m0);
            tma_load_2d_pair(st + 2 * P::kABytes + P::kBBytes, &tm_b_lo, bar, kb * kBK, brow);
          }
          }
        } else {
        mbar_expect_tx(&full_bar[s], tx);
        if (!MN) {
          const int kcol = k_begin + kb * kBK;
          tma_load_2d(st, &tm_a_hi, &full_bar[s], kcol, g.a_row0 + m0);
          tma_load_2d(st + 2 * P::kABytes, &tm_b_hi, &full_bar[s], kcol, g.b_row0 + n0);
          if (three) {
            tma_load_2d(st + P::kABytes, &tm_a_lo, &full_bar[s], kcol, g.a_row0 + m0);
            tma_load_2d(st + 2 * P::kABytes + P::kBBytes, &tm_b_lo, &full_bar[s], kcol, g.b_row0 + n0);
          }
        } else {
          constexpr int kBox = kBK * 128;  // bytes of one {32 mn, BK k} box
          const int krow = k_begin + kb * kBK;
          for (int i = 0; i < 128 / 32; ++i) {
            tma_load_2d(st + i * kBox, &tm_a_hi, &full_bar[s], m0 + 32 * i, g.a_row0 + krow);
            if (three) tma_load_2d(st + P::kABytes + i * kBox, &tm_a_lo, &full_bar[s], m0 + 32 * i, g.a_row0 + krow);
          }
          for (int i = 0; i < BN / 32; ++i) {
            tma_load_2d(st + 2 * P::kABytes + i * kBox, &tm_b_hi, &full_bar[s], n0 + 32 * i, g.b_row0 + krow);
            if (three)
              tma_load_2d(st + 2 * P::kABytes + P::kBBytes + i * kBox, &tm_b_lo, &full_bar[s], n0 + 32 * i,
                          g.b_row0 + krow);
          }
        }
        }
        }
        if (CG == 2) __syncwarp();  // lanes running ahead would spin on the next barrier and steal issue slots
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (CTA pair: the leader's warp runs the loop, one elected lane issues for both CTAs) =====
    if (rank == 0 && (CG == 2 || lane == 0)) {
      // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                             (static_cast<uint32_t>((128 * CG) >> 4) << 24) | (MN ? ((1u << 15) | (1u << 16)) : 0u);
      const int nch = n_kb < g.hi_chunks ? n_kb : g.hi_chunks;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % P::kStages;
        const uint32_t ph = (kb / P::kStages) & 1;
        const int chunk = 1 + (kb * nch) / n_kb;
        // the first MMA into an accumulator overwrites it (warp-uniform: derived from kb, not from per-lane state)
        const bool chunk_first = kb == 0 || chunk != 1 + ((kb - 1) * nch) / n_kb;
        const uint32_t d_hi = tmem_d + static_cast<uint32_t>(chunk * BN);
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (CG == 1 || elect_one()) {
        if (dbg && kb == 0) dbg[2] = clock64();
        const uint32_t st = smem_u32(smem + s * P::kStageBytes);
        constexpr uint32_t kLbo = kBK * 128;
        const uint64_t a_hi = MN ? make_mnmajor_desc(st, kLbo) : make_kmajor_desc(st);
        const uint64_t a_lo = MN ? make_mnmajor_desc(st + P::kABytes, kLbo) : make_kmajor_desc(st + P::kABytes);
        const uint64_t b_hi = MN ? make_mnmajor_desc(st + 2 * P::kABytes, kLbo) : make_kmajor_desc(st + 2 * P::kABytes);
        const uint64_t b_lo = MN ? make_mnmajor_desc(st + 2 * P::kABytes + P::kBBytes, kLbo)
                                 : make_kmajor_desc(st + 2 * P::kABytes + P::kBBytes);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          // K-major: 32 bytes per K=8 step inside the swizzle row; MN-major: one 1024-byte atom (8 k rows) per step
          const uint64_t adv = static_cast<uint64_t>((MN ? k * 1024 : k * 8 * 4) >> 4);
          const uint32_t acc_hi = (chunk_first && k == 0) ? 0u : 1u;
          const uint32_t acc_x = (kb == 0 && k == 0) ? 0u : 1u;
          if (CG == 2) {
            umma_tf32_pair(d_hi, a_hi + adv, b_hi + adv, idesc, acc_hi);
            if (three) {
              umma_tf32_pair(tmem_d, a_lo + adv, b_hi + adv, idesc, acc_x);
              umma_tf32_pair(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
            }
            continue;
          }
          umma_tf32(d_hi, a_hi + adv, b_hi + adv, idesc, acc_hi);
          if (three) {
            umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, acc_x);
            umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
          }
        }
        // frees the ring slot (in both CTAs of a pair) once these MMAs have read it
        if (CG == 2) umma_commit_pair(&empty_bar[s]);
        else umma_commit(&empty_bar[s]);
        }
        if (CG == 2) __syncwarp();
      }
      if (CG == 1 || elect_one()) {
        if (dbg) dbg[3] = clock64();
        if (CG == 2) umma_commit_pair(accum_bar);  // accumulators complete (each CTA drains its own 128 rows)
        else umma_commit(accum_bar);
      }
    }
  } else {
    // ===== epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q+32) = tile rows =====
    // Tiles move between shared and global memory by TMA only: a lane owns one output ROW, so direct global
    // accesses would touch 32 different rows per instruction (measured: the un-coalesced epilogue cost as much as
    // the mainloop). Each warp stages its 32 x (BN/2) sub-tile in 32-row x 32-float boxes (128-byte swizzle, bank
    // conflict free for one 16-byte chunk per lane) inside the operand ring, which is idle once accum_bar fired.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int we = warp - 2;                // epilogue warp index 0..7
    const int half = we >> 2;               // which half of the tile's columns this warp drains
    constexpr int kBoxes = BN / 64;         // 32-column boxes per warp and array
    constexpr int kBoxBytes = 32 * 128;
    const int cbase = half * (BN / 2);
    // first output row of this warp (weight-gradient form: c_row0 = row of the first partial slice of this launch)
    const int grow = g.c_row0 + (MN ? static_cast<int>(blockIdx.z) * g.M
                                    : (CG == 1 ? static_cast<int>(blockIdx.z) * g.c_zrows : 0)) + m0 + q * 32;
    unsigned char* stg = smem + we * (2 * kBoxes * kBoxBytes);  // [array 0 | array 1][box][32 rows x 128 B]
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    if (kPdlTrigger == 1 && warp == 2 && lane == 0) pdl_launch_dependents();  // this CTA is past its mainloop
    if (dbg && warp == 2 && lane == 0) dbg[4] = clock64();
    if (g.trace.buf && warp == 2 && lane == 0) tmem_slot[1] = static_cast<uint32_t>(trace_now());  // accumulators ready
    if (EPI == EPI_DGRAD_HIDDEN || EPI == EPI_DGRAD_ACCUM) {
      // the tile this epilogue combines with (saved layer output hi/lo, or the running state adjoint)
      constexpr int kArr = EPI == EPI_DGRAD_HIDDEN ? 2 : 1;
      constexpr int kFirstBox = kEarlyAux ? 1 : 0;  // box 0 arrived in aux_early long ago
      if (lane == 0) {
        mbar_expect_tx(&epi_bar[we], kArr * (kBoxes - kFirstBox) * kBoxBytes);
        const int xrow = g.x_row0 + m0 + q * 32;
#pragma unroll
        for (int b = kFirstBox; b < kBoxes; ++b) {
          tma_load_2d(stg + b * kBoxBytes, &tm.x0, &epi_bar[we], n0 + cbase + 32 * b, xrow);
          if (kArr == 2) tma_load_2d(stg + (kBoxes + b) * kBoxBytes, &tm.x1, &epi_bar[we], n0 + cbase + 32 * b, xrow);
        }
      }
      if (kEarlyAux) mbar_wait(&aux_bar[we], 0);  // box 0 is processed first; box 1 is awaited when it is reached
      else mbar_wait(&epi_bar[we], 0);
    }
    const int nch = n_kb < g.hi_chunks ? n_kb : g.hi_chunks;
    const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t r0[16], r1[16], r2[16], r3[16];
    // every accumulator's load for one 16-column chunk (results valid after tmem_ld_wait())
    auto issue_loads = [&](int cc_) {
      const int cl = cbase + cc_;
      tmem_ld16_async(lane_base + BN + cl, r0);  // hi*hi, first K chunk
      if (nch > 1) tmem_ld16_async(lane_base + 2 * BN + cl, r1);
      if (nch > 2) tmem_ld16_async(lane_base + 3 * BN + cl, r2);
      if (three) tmem_ld16_async(lane_base + cl, r3);  // cross terms
    };
#if HDPO_EPI_PIPE
    issue_loads(0);
#endif
#pragma unroll 1
    for (int cc = 0; cc < BN / 2; cc += 16) {
      const int c0 = cbase + cc;
      float v[16];
#if !HDPO_EPI_PIPE
      issue_loads(cc);
#endif
      tmem_ld_wait();
#if HDPO_EPI_PACK_SUM
      // (packed f32x2 adds: same order of additions as the scalar form, half the issued instructions)
      unsigned long long a2[8];
#pragma unroll
      for (int jp = 0; jp < 8; ++jp) a2[jp] = pack2(__uint_as_float(r0[2 * jp]), __uint_as_float(r0[2 * jp + 1]));
      if (nch > 1) {
#pragma unroll
        for (int jp = 0; jp < 8; ++jp)
          a2[jp] = add2_rn(a2[jp], pack2(__uint_as_float(r1[2 * jp]), __uint_as_float(r1[2 * jp + 1])));
      }
      if (nch > 2) {
#pragma unroll
        for (int jp = 0; jp < 8; ++jp)
          a2[jp] = add2_rn(a2[jp], pack2(__uint_as_float(r2[2 * jp]), __uint_as_float(r2[2 * jp + 1])));
      }
      if (three) {
#pragma unroll
        for (int jp = 0; jp < 8; ++jp)
          a2[jp] = add2_rn(a2[jp], pack2(__uint_as_float(r3[2 * jp]), __uint_as_float(r3[2 * jp + 1])));
      }
      if (EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_OUT) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cbase + cc + 4 * j4);
          a2[2 * j4] = add2_rn(a2[2 * j4], pack2(b4.x, b4.y));
          a2[2 * j4 + 1] = add2_rn(a2[2 * j4 + 1], pack2(b4.z, b4.w));
        }
      }
#pragma unroll
      for (int jp = 0; jp < 8; ++jp) unpack2(a2[jp], v[2 * jp], v[2 * jp + 1]);
#else
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = __uint_as_float(r0[j]);
        if (nch > 1) a += __uint_as_float(r1[j]);
        if (nch > 2) a += __uint_as_float(r2[j]);
        if (three) a += __uint_as_float(r3[j]);
        v[j] = a;
      }
      if (EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_OUT) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cbase + cc + 4 * j4);
          v[4 * j4 + 0] += b4.x;
          v[4 * j4 + 1] += b4.y;
          v[4 * j4 + 2] += b4.z;
          v[4 * j4 + 3] += b4.w;
        }
      }
#endif
#if HDPO_EPI_PIPE
      // software pipeline: the accumulators of this chunk are summed into v; the next chunk's TMEM loads fly while the
      // activation / split / staging of this one runs
      if (cc + 16 < BN / 2) issue_loads(cc + 16);
#endif
      const int n = n0 + c0;
      if (kEarlyAux && cc == 32) mbar_wait(&epi_bar[we], 0);  // second box of the combined tile has landed by now
      // staging addresses of this lane's four 16-byte chunks (array 0; array 1 is kBoxes boxes further)
      unsigned char* box = stg + (cc >> 5) * kBoxBytes + lane * 128;
      const int ch0 = (cc & 31) >> 2;
      auto chunk_ptr = [&](int arr, int j4) {
        return reinterpret_cast<float4*>(box + arr * (kBoxes * kBoxBytes) + (((ch0 + j4) ^ (lane & 7)) << 4));
      };
      // where the combined tile's chunks are read from: the early box lives in its own region
      auto aux_ptr = [&](int arr, int j4) {
        if (kEarlyAux && cc < 32)
          return reinterpret_cast<const float4*>(aux_early + we * (2 * kBoxBytes) + arr * kBoxBytes + lane * 128 +
                                                 (((ch0 + j4) ^ (lane & 7)) << 4));
        return reinterpret_cast<const float4*>(chunk_ptr(arr, j4));
      };
      if (EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_OUT) {
        if (g.act == HDPO_ACT_ELU) {
          elu_inplace<16, HDPO_EPI_PACK_ELU != 0>(v);  // branch-free, element chains overlap
        } else {
          dispatch_act(g.act, [&](auto tag) {
            constexpr int ACT = decltype(tag)::value;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_fwd_t<ACT>(v[j]);
          });
        }
      } else if (EPI == EPI_DGRAD_HIDDEN) {
        float h[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 a = *aux_ptr(0, j4);
          const float4 b = *aux_ptr(1, j4);
          h[4 * j4 + 0] = a.x + b.x;
          h[4 * j4 + 1] = a.y + b.y;
          h[4 * j4 + 2] = a.z + b.z;
          h[4 * j4 + 3] = a.w + b.w;
        }
        dispatch_act(g.act, [&](auto tag) {
          constexpr int ACT = decltype(tag)::value;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= act_grad_out_t<ACT>(h[j]);
        });
      } else if (EPI == EPI_DGRAD_ACCUM) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 c = *chunk_ptr(0, j4);
          v[4 * j4 + 0] += c.x;
          v[4 * j4 + 1] += c.y;
          v[4 * j4 + 2] += c.z;
          v[4 * j4 + 3] += c.w;
        }
      }
      if (EPI == EPI_DGRAD_HIDDEN) {
        if (g.colsum_part) {
          // bias-gradient partials: column sums over this warp's 32 rows by a reduce-scatter butterfly (16 shuffles
          // for 16 columns); lane pairs end up with column 8*b4 + 4*b3 + 2*b2 + b1 of the chunk (b_i = lane bits)
          float w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = v[j];
#pragma unroll
          for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < h; ++j) {
              const float send = up ? w[j] : w[j + h];
              const float recv = __shfl_xor_sync(0xffffffffu, send, o);
              w[j] = (up ? w[j + h] : w[j]) + recv;
            }
          }
          w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
          if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            const size_t rblk = static_cast<size_t>(grow) >> 5;
            g.colsum_part[rblk * g.ldc + n + col] = w[0];
          }
        }
      }
      if (EPI == EPI_FWD_HIDDEN || EPI == EPI_DGRAD_HIDDEN) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float4 hi, lo;
          hi.x = tf32_hi(v[4 * j4 + 0]);
          hi.y = tf32_hi(v[4 * j4 + 1]);
          hi.z = tf32_hi(v[4 * j4 + 2]);
          hi.w = tf32_hi(v[4 * j4 + 3]);
          // the remainder is rounded (not left to the tensor core's truncation) so that its error is unbiased;
          // v - hi as one packed FMA per pair (hi * -1 + v: exact, like the subtraction)
#if HDPO_EPI_PACK_SUM
          const unsigned long long m1 = pack2(-1.f, -1.f);
          float d0, d1, d2, d3;
          unpack2(fma2_rn(pack2(hi.x, hi.y), m1, pack2(v[4 * j4 + 0], v[4 * j4 + 1])), d0, d1);
          unpack2(fma2_rn(pack2(hi.z, hi.w), m1, pack2(v[4 * j4 + 2], v[4 * j4 + 3])), d2, d3);
#else
          const float d0 = v[4 * j4 + 0] - hi.x, d1 = v[4 * j4 + 1] - hi.y, d2 = v[4 * j4 + 2] - hi.z,
                      d3 = v[4 * j4 + 3] - hi.w;
#endif
          lo.x = tf32_hi(d0);
          lo.y = tf32_hi(d1);
          lo.z = tf32_hi(d2);
          lo.w = tf32_hi(d3);
          *chunk_ptr(0, j4) = hi;
          *chunk_ptr(1, j4) = lo;
        }
      } else {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          *chunk_ptr(0, j4) = make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      }
      // staged box -> global as soon as its 32 columns are complete (one TMA store per 32 x 32 box and array): the
      // store of box b drains while box b + 1 is still being computed
      if ((cc & 31) == 16) {
        const int b = cc >> 5;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm.c0, stg + b * kBoxBytes, n0 + cbase + 32 * b, grow);
          if (EPI == EPI_FWD_HIDDEN || EPI == EPI_DGRAD_HIDDEN)
            tma_store_2d(&tm.c1, stg + (kBoxes + b) * kBoxBytes, n0 + cbase + 32 * b, grow);
          tma_store_commit();
        }
      }
    }
    if (dbg && warp == 2 && lane == 0) dbg[5] = clock64();
    if (g.trace.buf && warp == 2 && lane == 0) tmem_slot[2] = static_cast<uint32_t>(trace_now());  // tile staged
    if (lane == 0) {
      tma_store_wait_read();  // shared memory (and TMEM) are released right after the final barrier
      if (dbg && warp == 2) dbg[6] = clock64();
      if (g.trace.buf && warp == 2) tmem_slot[3] = static_cast<uint32_t>(trace_now());  // staged tile read by TMA
    }
  }
  if (kPdlTrigger == 2 && threadIdx.x == 64) pdl_launch_dependents();  // first epilogue thread: its tile is on its way
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // neither CTA may release shared memory / TMEM the pair's MMAs and commits still use
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_pair(tmem_d, kAccums * BN);
    else tmem_dealloc(tmem_d, kAccums * BN);
  }
  if (dbg && threadIdx.x == 0) dbg[7] = clock64();
  if (threadIdx.x == 0 && g.trace.buf)
    trace_emit(g.trace, trace_t0, blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z),
               // aux: ns/32 from start to "accumulators ready" | "tile staged" | "stores read" (10 bits each)
               (((tmem_slot[1] - static_cast<uint32_t>(trace_t0)) >> 5) & 1023u) |
                   ((((tmem_slot[2] - static_cast<uint32_t>(trace_t0)) >> 5) & 1023u) << 10) |
                   ((((tmem_slot[3] - static_cast<uint32_t>(trace_t0)) >> 5) & 1023u) << 20));
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HDPO_TC_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v;
}

int make_tensor_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  HDPO_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  HDPO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0, "TMA needs 16-byte aligned rows");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
#ifdef HDPO_TMA_L2_128  // (A/B: 256-byte promotion measured 0.3 % faster on the wide step, tools/epi_ab.sh)
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
#else
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
#endif
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", static_cast<int>(r),
              static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
              static_cast<unsigned long long>(ld), box_rows);
    return HDPO_E_CUDA;
  }
  return HDPO_OK;
}

template <int BN, int EPI, bool MN = false, int CG = 1, int OCC = 1>
static int launch(const GemmTcMaps& tm, const GemmTcArgs& g_in, void* stream) {
  GemmTcArgs g = g_in;
  if (g.hi_chunks <= 0 || g.hi_chunks > kHiChunks) g.hi_chunks = kHiChunks;
  {
    // HDPO_TC_HI_CHUNKS (accuracy experiments, tools/wg_accuracy.py): accumulators the hi*hi term of the K-major GEMMs is
    // spread over (default 3; 1 = what a 256 x 256 pair tile could afford in TMEM)
    static int env_chunks = -1;
    if (env_chunks < 0) {
      const char* e = getenv("HDPO_TC_HI_CHUNKS");
      env_chunks = e ? atoi(e) : 0;
    }
    if (!MN && env_chunks >= 1 && env_chunks <= kHiChunks) g.hi_chunks = env_chunks;
  }
  g.trace = trace_ref((g_in.trace.tag << 8) | (static_cast<unsigned>(EPI) << 4) | (MN ? 8u : 0u) | (BN == 128 ? 1u : 0u));
  auto k = gemm_tc_kernel<BN, EPI, MN, CG, OCC>;
  static bool configured = false;
  if (!configured) {
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemPlan<BN, CG, false, OCC>::kTotal));
    if (OCC == 2)  // two CTAs of ~82 / ~98 KB each must fit next to each other
      HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  const unsigned nz = (MN || (CG == 1 && g.k_per_split > 0)) ? static_cast<unsigned>(g.K / g.k_per_split) : 1u;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = CG == 2 ? (MN ? dim3(2, (g.M / 256) * (g.N / BN), static_cast<unsigned>(g.K / g.k_per_split))
                              : dim3(2, g.M / 256, g.N / BN))
                        : dim3(g.N / BN, g.M / 128, nz);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SmemPlan<BN, CG, false, OCC>::kTotal;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: prologue overlaps the previous kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;                 // CTA pair
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 2 : 1;
  HDPO_CUDA_OK(cudaLaunchKernelEx(&cfg, k, tm, g));
  count_launch();
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

template <int BN>
static int launch_epi(const GemmTcMaps& tm, const GemmTcArgs& g, int epi, void* stream) {
  switch (epi) {
    case EPI_FWD_HIDDEN: return launch<BN, EPI_FWD_HIDDEN>(tm, g, stream);
    case EPI_FWD_OUT: return launch<BN, EPI_FWD_OUT>(tm, g, stream);
    case EPI_DGRAD_HIDDEN: return launch<BN, EPI_DGRAD_HIDDEN>(tm, g, stream);
    case EPI_DGRAD_ACCUM: return launch<BN, EPI_DGRAD_ACCUM>(tm, g, stream);
    case EPI_STORE: return launch<BN, EPI_STORE>(tm, g, stream);
  }
  set_error("bad epilogue %d", epi);
  return HDPO_E_INVALID;
}

int gemm(const GemmTcMaps& tm, const GemmTcArgs& g, int epi, int bn, void* stream) {
  if (bn > kBnMulti) return wp::gemm_multi(tm, g, epi, bn - kBnMulti, stream);
  const int bn_cols = bn == kBnPair ? 128 : ((bn == kBnPair64 || bn == kBn64x2) ? 64 : bn);
  HDPO_REQUIRE(g.M % 128 == 0 && g.N % bn_cols == 0 && g.K % kBK == 0 && g.K > 0,
               "tcgen05 GEMM shape %dx%dx%d not tileable", g.M, g.N, g.K);
  HDPO_REQUIRE(g.n_pass == 1 || g.n_pass == 3, "n_pass must be 1 or 3");
  if (g.k_per_split > 0) {  // split-K of a K-major GEMM: single-CTA tiles, plain / bias epilogues (partials are summed later)
    HDPO_REQUIRE((bn == 64 || bn == 128) && (epi == EPI_STORE || epi == EPI_FWD_OUT) && g.k_per_split % kBK == 0 &&
                     g.K % g.k_per_split == 0 && g.c_zrows >= g.M,
                 "split-K GEMM %dx%dx%d (k_per_split %d, bn %d, epilogue %d) not supported", g.M, g.N, g.K, g.k_per_split,
                 bn, epi);
  }
  if (bn == kBnPair) {  // CTA-pair form: 256 x 128 tiles (the B map must have BN / 2 = 64-row boxes)
    HDPO_REQUIRE(g.M % 256 == 0 && g.N % 128 == 0, "CTA-pair GEMM shape %dx%d not tileable", g.M, g.N);
    switch (epi) {
      case EPI_FWD_HIDDEN: return launch<128, EPI_FWD_HIDDEN, false, 2>(tm, g, stream);
      case EPI_DGRAD_HIDDEN: return launch<128, EPI_DGRAD_HIDDEN, false, 2>(tm, g, stream);
      case EPI_STORE: return launch<128, EPI_STORE, false, 2>(tm, g, stream);
    }
    set_error("the CTA-pair GEMM has no epilogue %d", epi);
    return HDPO_E_INVALID;
  }
  if (bn == kBnPair64) {  // 256 x 64 pair tiles, two CTAs per SM (the B map must have 32-row boxes)
    HDPO_REQUIRE(g.M % 256 == 0, "CTA-pair GEMM shape %dx%d not tileable", g.M, g.N);
    switch (epi) {
      case EPI_FWD_HIDDEN: return launch<64, EPI_FWD_HIDDEN, false, 2, 2>(tm, g, stream);
      case EPI_DGRAD_HIDDEN: return launch<64, EPI_DGRAD_HIDDEN, false, 2, 2>(tm, g, stream);
    }
    set_error("the 256 x 64 CTA-pair GEMM has no epilogue %d", epi);
    return HDPO_E_INVALID;
  }
  if (bn == kBn64x2) {  // 128 x 64 single-CTA tiles, two CTAs per SM
    switch (epi) {
      case EPI_FWD_HIDDEN: return launch<64, EPI_FWD_HIDDEN, false, 1, 2>(tm, g, stream);
      case EPI_FWD_OUT: return launch<64, EPI_FWD_OUT, false, 1, 2>(tm, g, stream);
      case EPI_DGRAD_HIDDEN: return launch<64, EPI_DGRAD_HIDDEN, false, 1, 2>(tm, g, stream);
      case EPI_DGRAD_ACCUM: return launch<64, EPI_DGRAD_ACCUM, false, 1, 2>(tm, g, stream);
    }
    set_error("the two-per-SM GEMM has no epilogue %d", epi);
    return HDPO_E_INVALID;
  }
  if (bn == 128) return launch_epi<128>(tm, g, epi, stream);
  if (bn == 64) return launch_epi<64>(tm, g, epi, stream);
  set_error("unsupported BN %d", bn);
  return HDPO_E_INVALID;
}

int gemm_wgrad(const GemmTcMaps& tm, const GemmTcArgs& g, int bn, void* stream) {
  HDPO_REQUIRE(g.M % 128 == 0 && g.N % (bn == kBnPair ? 128 : (bn == kBnPair64 ? 64 : bn)) == 0 &&
                   g.k_per_split % kBK == 0 && g.k_per_split > 0 &&
                   g.K % g.k_per_split == 0,
               "tcgen05 weight-gradient GEMM shape %dx%dx%d (k_per_split %d) not tileable", g.M, g.N, g.K, g.k_per_split);
  HDPO_REQUIRE(g.n_pass == 1 || g.n_pass == 3, "n_pass must be 1 or 3");
  if (bn == kBnPair) {  // 256 x 128 CTA-pair tiles
    HDPO_REQUIRE(g.M % 256 == 0 && g.N % 128 == 0, "CTA-pair weight-gradient shape %dx%d not tileable", g.M, g.N);
    return launch<128, EPI_STORE, true, 2>(tm, g, stream);
  }
  if (bn == kBnPair64) {  // 256 x 64 CTA-pair tiles (one CTA per SM: the 4-stage ring of the single-CTA form)
    HDPO_REQUIRE(g.M % 256 == 0 && g.N % 64 == 0, "CTA-pair weight-gradient shape %dx%d not tileable", g.M, g.N);
    return launch<64, EPI_STORE, true, 2>(tm, g, stream);
  }
  if (bn == 128) return launch<128, EPI_STORE, true>(tm, g, stream);
  if (bn == 64) return launch<64, EPI_STORE, true>(tm, g, stream);
  set_error("unsupported BN %d", bn);
  return HDPO_E_INVALID;
}

// split fp32 -> (tf32 hi, exact remainder lo)
__global__ void __launch_bounds__(256) split_hi_lo_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                          float* __restrict__ lo, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float h = tf32_hi(v);
  hi[i] = h;
  lo[i] = tf32_hi(v - h);
}

}  // namespace tc
}  // namespace hdpo

using namespace hdpo;

// Test hook: C[M,N] = A[M,K] * B[N,K]^T on the tcgen05 path (n_pass = 3: 3xTF32, 1: TF32). A, B, C dense row-major
// device arrays with M % 128 == 0, N % 64 == 0, K % 32 == 0; scratch = 2*(M*K + N*K) floats of device memory.
extern "C" int hdpo_debug_gemm_tc(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                                  int32_t n_pass, float* scratch, void* stream) {
  HDPO_REQUIRE(A && B && C && scratch, "null argument");
  HDPO_REQUIRE(M % 128 == 0 && N % 64 == 0 && K % 32 == 0 && M > 0 && N > 0 && K > 0, "shape not tileable");
  float* a_hi = scratch;
  float* a_lo = a_hi + static_cast<size_t>(M) * K;
  float* b_hi = a_lo + static_cast<size_t>(M) * K;
  float* b_lo = b_hi + static_cast<size_t>(N) * K;
  const size_t na = static_cast<size_t>(M) * K, nb = static_cast<size_t>(N) * K;
  tc::split_hi_lo_kernel<<<static_cast<unsigned>((na + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, a_hi, a_lo, na);
  tc::split_hi_lo_kernel<<<static_cast<unsigned>((nb + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(B, b_hi, b_lo, nb);
  count_launch();
  count_launch();
  HDPO_LAUNCH_OK();
  const int bn_multi = wp::pick_bn_multi(M, N);  // hdpo_debug_set_tc_multi(1): the multi-tile CTA-pair form
  const int bn = bn_multi ? bn_multi : tc::pick_bn_pair(M, N);
  tc::GemmTcMaps tm{};
  int rc;
  if ((rc = tc::make_tensor_map(&tm.a_hi, a_hi, M, K, K, 128))) return rc;
  if ((rc = tc::make_tensor_map(&tm.a_lo, a_lo, M, K, K, 128))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_hi, b_hi, N, K, K, tc::b_box_rows(bn)))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_lo, b_lo, N, K, K, tc::b_box_rows(bn)))) return rc;
  if ((rc = tc::make_tensor_map(&tm.c0, C, M, N, N, tc::kBoxRowsC))) return rc;
  tm.c1 = tm.x0 = tm.x1 = tm.c0;
  tc::GemmTcArgs g{};
  g.M = M;
  g.N = N;
  g.K = K;
  g.n_pass = n_pass;
  g.ldc = N;
  return tc::gemm(tm, g, tc::EPI_STORE, bn, stream);
}

// Same as hdpo_debug_gemm_tc with per-CTA clock64 stamps: dbg_clock[8 * n_ctas] (device), see gemm_tc_kernel.
// epi = 4 (plain store) or 0 (bias + ELU + (hi, lo) split; C_lo and bias must then be given)
extern "C" int hdpo_debug_gemm_tc_timeline(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                                           int32_t n_pass, float* scratch, long long* dbg_clock, void* stream,
                                           int32_t epi, float* C_lo, const float* bias) {
  HDPO_REQUIRE(A && B && C && scratch && dbg_clock, "null argument");
  HDPO_REQUIRE(M % 128 == 0 && N % 64 == 0 && K % 32 == 0 && M > 0 && N > 0 && K > 0, "shape not tileable");
  float* a_hi = scratch;
  float* a_lo = a_hi + static_cast<size_t>(M) * K;
  float* b_hi = a_lo + static_cast<size_t>(M) * K;
  float* b_lo = b_hi + static_cast<size_t>(N) * K;
  const int bn_multi = wp::pick_bn_multi(M, N);
  const int bn = bn_multi ? bn_multi : tc::pick_bn_pair(M, N);
  tc::GemmTcMaps tm{};
  int rc;
  if ((rc = tc::make_tensor_map(&tm.a_hi, a_hi, M, K, K, 128))) return rc;
  if ((rc = tc::make_tensor_map(&tm.a_lo, a_lo, M, K, K, 128))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_hi, b_hi, N, K, K, tc::b_box_rows(bn)))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_lo, b_lo, N, K, K, tc::b_box_rows(bn)))) return rc;
  if ((rc = tc::make_tensor_map(&tm.c0, C, M, N, N, tc::kBoxRowsC))) return rc;
  tm.c1 = tm.x0 = tm.x1 = tm.c0;
  tc::GemmTcArgs g{};
  g.M = M;
  g.N = N;
  g.K = K;
  g.n_pass = n_pass;
  g.ldc = N;
  g.dbg_clock = dbg_clock;
  if (epi == tc::EPI_FWD_HIDDEN) {
    HDPO_REQUIRE(C_lo && bias, "null argument");
    if ((rc = tc::make_tensor_map(&tm.c1, C_lo, M, N, N, tc::kBoxRowsC))) return rc;
    g.bias = bias;
    g.act = HDPO_ACT_ELU;
    return tc::gemm(tm, g, tc::EPI_FWD_HIDDEN, bn, stream);
  }
  return tc::gemm(tm, g, tc::EPI_STORE, bn, stream);
}

// sum of the K-slice partials: C[i] = sum_z part[z*slice + i] (double accumulation)
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int nz, size_t slice,
                                                           float* __restrict__ C) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= slice) return;
  double s = 0.0;
  for (int z = 0; z < nz; ++z) s += static_cast<double>(part[z * slice + i]);
  C[i] = static_cast<float>(s);
}

// Test hook (weight-gradient form): C[M,N] = A[K,M]^T * B[K,N] with A, B row-major [K][M] / [K][N] device arrays,
// M % 128 == 0, N % 64 == 0, K % k_per_split == 0, k_per_split % 32 == 0.
// scratch = 2*(K*M + K*N) + (K/k_per_split)*M*N floats.
extern "C" int hdpo_debug_gemm_tc_wgrad(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                                        int32_t k_per_split, int32_t n_pass, float* scratch, void* stream) {
  HDPO_REQUIRE(A && B && C && scratch, "null argument");
  HDPO_REQUIRE(M % 128 == 0 && N % 64 == 0 && k_per_split > 0 && k_per_split % 32 == 0 && K % k_per_split == 0,
               "shape not tileable");
  const size_t na = static_cast<size_t>(K) * M, nb = static_cast<size_t>(K) * N;
  float* a_hi = scratch;
  float* a_lo = a_hi + na;
  float* b_hi = a_lo + na;
  float* b_lo = b_hi + nb;
  float* part = b_lo + nb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tc::split_hi_lo_kernel<<<static_cast<unsigned>((na + 255) / 256), 256, 0, st>>>(A, a_hi, a_lo, na);
  tc::split_hi_lo_kernel<<<static_cast<unsigned>((nb + 255) / 256), 256, 0, st>>>(B, b_hi, b_lo, nb);
  count_launch();
  count_launch();
  HDPO_LAUNCH_OK();
  const char* wg_pair = getenv("HDPO_WG_PAIR");
  const bool wg_pair_on = !wg_pair || atoi(wg_pair) != 0;  // (same default as the rollout: pairs where the shape allows)
  const int bn = (wg_pair_on && M % 256 == 0) ? (N % 128 == 0 ? tc::kBnPair : tc::kBnPair64) : tc::pick_bn(N);
  tc::GemmTcMaps tm{};
  int rc;
  if ((rc = tc::make_tensor_map(&tm.a_hi, a_hi, K, M, M, 32, true))) return rc;
  if ((rc = tc::make_tensor_map(&tm.a_lo, a_lo, K, M, M, 32, true))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_hi, b_hi, K, N, N, 32, true))) return rc;
  if ((rc = tc::make_tensor_map(&tm.b_lo, b_lo, K, N, N, 32, true))) return rc;
  if ((rc = tc::make_tensor_map(&tm.c0, part, static_cast<uint64_t>(K / k_per_split) * M, N, N, tc::kBoxRowsC))) return rc;
  tm.c1 = tm.x0 = tm.x1 = tm.c0;
  tc::GemmTcArgs g{};
  g.M = M;
  g.N = N;
  g.K = K;
  g.k_per_split = k_per_split;
  g.c_slice = static_cast<size_t>(M) * N;
  g.n_pass = n_pass;
  g.ldc = N;
  if ((rc = tc::gemm_wgrad(tm, g, bn, stream))) return rc;
  const size_t slice = static_cast<size_t>(M) * N;
  sum_partials_kernel<<<static_cast<unsigned>((slice + 255) / 256), 256, 0, st>>>(part, K / k_per_split, slice, C);
  count_launch();
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

#else  // HDPO_EMU: tensor cores cannot be emulated; the entry point exists and refuses.

extern "C" int hdpo_debug_gemm_tc_wgrad(const float*, const float*, float*, int32_t, int32_t, int32_t, int32_t, int32_t,
                                        float*, void*) {
  hdpo::set_error("tcgen05 path is not available in the host-thread emulator");
  return HDPO_E_INVALID;
}

extern "C" int hdpo_debug_gemm_tc(const float*, const float*, float*, int32_t, int32_t, int32_t, int32_t, float*, void*) {
  hdpo::set_error("tcgen05 path is not available in the host-thread emulator");
  return HDPO_E_INVALID;
}

extern "C" int hdpo_debug_gemm_tc_timeline(const float*, const float*, float*, int32_t, int32_t, int32_t, int32_t, float*,
                                           long long*, void*, int32_t, float*, const float*) {
  hdpo::set_error("tcgen05 path is not available in the host-thread emulator");
  return HDPO_E_INVALID;
}

#endif
