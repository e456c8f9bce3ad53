// wide_persist.cuh - persistent forward / adjoint sweeps of the wide-net rollout (VanillaWarehouse) on tcgen05.
//
// ONE kernel launch runs all T periods of a direction: every CTA pair (cluster of 2, tcgen05 cta_group::2) walks a
// host-built list of 256-row x 128/64-column GEMM tiles (all layers of all periods of the scenario row tiles it
// serves) without leaving the SM. Tiles of consecutive layers exchange their operands through the L2-resident
// activation tapes (TMA store -> release flag -> acquire -> TMA load); the policy head + simulator period run on
// dedicated head warps of the same kernel (one thread per scenario, transposed state so every access is coalesced).
// See wide_persist.cu for the pipeline (TMA ring, double-buffered TMEM accumulators, register-accumulating epilogue).
#pragma once

#include "hdpo_internal.cuh"

#ifndef HDPO_EMU
#include "gemm_tc.cuh"

namespace hdpo {
namespace wp {

// Multi-tile CTA-pair GEMM behind the interface of tc::gemm (same maps and arguments; the B map needs bn / 2-row
// boxes): one launch per layer, every pair walks several 256 x bn tiles with the epilogue of one tile overlapping the
// MMAs of the next (see wide_persist.cu). tc::gemm() routes the selectors tc::kBnMulti + 64 | 128 here.
int gemm_multi(const tc::GemmTcMaps& tm, const tc::GemmTcArgs& g, int epi, int bn, void* stream);
// Routing: fewest 256 x bn tiles of a launch that goes to gemm_multi (below that the single-tile forms spread the work
// over more SMs). Opt-in: HDPO_TC_MULTI = 1 enables it, HDPO_TC_MULTI_MIN sets the threshold (default 48);
// set_multi_min_tiles overrides both (> 0 threshold, 0 never, < 0 back to the default).
int multi_min_tiles();
void set_multi_min_tiles(int min_tiles);
// tc::gemm selector (tc::kBnMulti + 64 | 128) when a rows x cols GEMM is routed to the multi-tile form, else 0
inline int pick_bn_multi(int rows, int cols) {
  if (rows % 256 != 0 || cols % 64 != 0) return 0;
  const int bn = cols % 128 == 0 ? 128 : 64;
  return (rows / 256) * (cols / bn) >= multi_min_tiles() ? tc::kBnMulti + bn : 0;
}

constexpr int kRowTile = 256;  // scenarios per row tile (one CTA pair, M = 256)
constexpr int kMaxW = 4;       // warehouses the per-thread head handles

// Everything the persistent sweeps need, filled by rollout_wide.cu from its workspace plan (device pointers).
struct Ctx {
  // shapes
  int B, Bp, T, n;                  // scenarios, padded to kRowTile, periods, linear layers
  int w[HDPO_MAX_LAYERS + 1];       // true widths
  int wp[HDPO_MAX_LAYERS + 1];      // widths padded to 64
  int act[HDPO_MAX_LAYERS];         // activation after layer l
  int save, n_pass;
  int period_shift, ignore_periods, t_stride, demand_layout, B_total;
  // problem
  int S, W, L, Lw;
  int lost, profit, has_edge, transshipment, discrete;
  float wub;
  const int32_t* adjacency;  // [W][S] or null
  // packed parameters
  const float *W_hi[HDPO_MAX_LAYERS], *W_lo[HDPO_MAX_LAYERS], *bias[HDPO_MAX_LAYERS];      // [wp[l+1]][wp[l]]
  const float *WT_hi[HDPO_MAX_LAYERS], *WT_lo[HDPO_MAX_LAYERS];                            // [wp[l]][wp[l+1]]
  // tapes (row-major [slots][Bp][width]); slots = T when save else 1 (X: T+1 / 2)
  float *X, *X_hi, *X_lo;
  float *act_hi[HDPO_MAX_LAYERS], *act_lo[HDPO_MAX_LAYERS];  // last layer: act_hi = fp32 output, act_lo = null
  float *gz_hi[HDPO_MAX_LAYERS], *gz_lo[HDPO_MAX_LAYERS];
  float* gX;                        // [Bp][wp[0]] state adjoint (adjoint sweep)
  float* csum[HDPO_MAX_LAYERS];     // [T*Bp/32][wp[l+1]] column sums of 32-row blocks of gz_l (hidden l)
  // inputs / outputs of the rollout
  const float* demands;
  HdpoStatics st;
  float *cost_b, *report_b, *reward_tb;
  // private scratch of the persistent path (extra_bytes())
  void* extra;
  void* stream;
};

bool enabled();                                  // HDPO_WIDE_PERSIST=1 (default off) or set_enabled()
void set_enabled(int on);
bool eligible(const HdpoRolloutDesc* d);         // shapes the persistent kernels handle
size_t extra_bytes(const HdpoRolloutDesc* d, int Bp, const int* wp, int n);
void set_trace(unsigned long long* buf, int cap_per_role);  // debug: [74 pairs][4 roles][cap] x {tag, ns}
int forward(const Ctx& c);                       // all T periods; the initial state must be in X[0], X_hi/lo[0]
int backward(const Ctx& c, float g_total, float g_report);  // adjoint sweep: fills gz tapes + csum (weight gradients follow)

}  // namespace wp
}  // namespace hdpo
#endif
