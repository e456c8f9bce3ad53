// rollout_sym.cuh - SymmetryAware policy head (context / store / warehouse nets, weight-duplicated
// store net) for the wide rollout. The context net and the context half of the first store / warehouse layers run as
// tile GEMMs in rollout_wide.cu (the "trunk": context layers + one 64-column projection); this unit holds the
// per-scenario part: the store net's local layers for every store, the warehouse net, proportional allocation
// (neural_networks.py:111-138, old epsilon 1e-15), the simulator period (environment.py:110-270) and the adjoint of
// all of it, with the local-net weight gradients accumulated in registers / shared memory.
#pragma once

#include "hdpo_internal.cuh"

namespace hdpo {
namespace sym {

constexpr int kMaxHH = 2;      // HxH hidden layers of the store / warehouse nets (n_layers - 2)
constexpr int kMaxBlocks = 16; // parameter blocks of the two local nets

// one parameter block of the local nets: where it sits in a per-warp gradient slab and in the flat parameter vector
struct Block {
  int q, q_ld;       // slab offset, slab row stride
  int rows, cols;    // true extent
  int dst, dst_ld;   // flat parameter offset of element (0,0), row stride there
};

struct Cfg {
  int B, Bp, S, SP, L, Lw, C;
  int ldx, ldy, ldo;  // leading dimensions: state row, projection row (64), store-output tape row
  int nS;             // S * L
  // store net: local inputs (L pipeline slots, mean, std, underage, lead time), hidden widths, activations
  int s_in, s_in4, s_xs, s_kq, s_nhh, s_hact, s_oact, s_w[kMaxHH + 1];
  int s_ld0;          // row stride of the store net's first weight in the flat vector: L + 4 + C
  int g_s_w0, g_s_b0, g_s_wh[kMaxHH], g_s_bh[kMaxHH], g_s_wo, g_s_bo;
  // warehouse net: local inputs (Lw pipeline slots)
  int w_nhh, w_hact, w_oact, w_w[kMaxHH + 1];
  int w_ld0;          // Lw + C
  int g_w_w0, g_w_b0, g_w_wh[kMaxHH], g_w_bh[kMaxHH], g_w_wo, g_w_bo;
  int P;              // parameters of all three nets
  // shared-memory weight block (floats)
  int m_s_wt0, m_s_wth[kMaxHH], m_s_bh[kMaxHH], m_s_wo, m_s_bo;
  int m_w_wt0, m_w_wth[kMaxHH], m_w_bh[kMaxHH], m_w_wo, m_w_bo, m_total, m_total_simt;
  // tensor-core mode (tc): pre-split (hi, lo) weight copies W[n][k] for the mma.sync fragments (mma32.cuh): first
  // store layer (local columns, row stride s_xs), hidden layers as W[n][k] and transposed W^T[k][n] (row stride 36)
  int tc, tc_fwd, tc_recompute;  // adjoint dgrad / wgrad; forward head; recompute inside the adjoint head
  int m_s_w0n_hi, m_s_w0n_lo, m_s_whn_hi[kMaxHH], m_s_whn_lo[kMaxHH], m_s_whk_hi[kMaxHH], m_s_whk_lo[kMaxHH];
  // per-warp gradient slab (floats) and its blocks
  int q_s_w0, q_s_wh[kMaxHH], q_s_bh[kMaxHH], q_s_wo, q_s_bo;
  int q_w_w0, q_w_wh[kMaxHH], q_w_bh[kMaxHH], q_w_wo, q_w_bo, q_total;
  int n_blocks;
  Block blk[kMaxBlocks];
  // simulator
  int lost, profit, has_edge, discrete;
  int t_stride, demand_layout, demand_bstride;
  float wub, eps;
};

struct PeriodArgs {
  int tt;         // demand column (t + period_shift)
  int in_report;  // t >= ignore_periods
  const float* demands;
  HdpoStatics st;
};

bool supported(const HdpoRolloutDesc* d);
// ldx / ldy: leading dimensions the trunk uses for the state rows and the projection rows (ldy must be 64)
int build_cfg(const HdpoRolloutDesc* d, int B, int Bp, int ldx, int ldy, Cfg* c);
// warps of the adjoint head (its gradient slabs are [bwd_warps][q_total] floats); launch shape of both heads
int bwd_warps(const Cfg& c);
int fwd_smem_floats_per_warp(const Cfg& c);
int bwd_smem_floats_per_warp(const Cfg& c);

// trunk projection layer: rows 0..31 = context columns of the store net's first layer (+ its bias), rows 32..63 the
// warehouse net's; hi/lo/transposed copies as pack_layer_kernel makes them (null in fp32 mode)
int pack_projection(const Cfg& c, const float* params, int Kp, float* Wp, float* bp, float* W_lo, float* WT, float* WT_lo,
                    void* stream);

// forward head of one period: X (state rows), PRJ (projection rows) -> Xn (+ its tf32 split), costs, store outputs
int head_fwd(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ, float* Xn,
             float* Xn_hi, float* Xn_lo, float* so_tape, float* cost_b, float* report_b, float* reward_t, void* stream);
// adjoint head: gX in/out as in the vanilla head; gPRJ (hi half, lo half or null); slabs accumulate over the periods
int head_bwd(const Cfg& c, const PeriodArgs& a, const float* params, const float* X, const float* PRJ,
             const float* so_tape, float* gX, float* gPRJ, float* gPRJ_lo, float rb, float* slabs, int first,
             void* stream);
// grad[local-net parameters] = fixed-order sum of the slabs
int reduce_slabs(const Cfg& c, const float* slabs, float* grad, void* stream);

}  // namespace sym
}  // namespace hdpo
