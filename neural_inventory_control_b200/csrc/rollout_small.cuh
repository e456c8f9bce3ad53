// rollout_small.cuh - shared declarations of the fused small-net rollout path (K1/K2 for the one-store and
// serial policies: hidden width <= 32, trunk of <= 4 hidden layers).
#pragma once

#include "hdpo_internal.cuh"

namespace hdpo {
namespace small {

constexpr int H = 32;        // padded hidden width
constexpr int HS = 36;       // shared-memory row stride of a hidden vector (16B aligned, conflict-free float4 rows)
constexpr int kMaxOut = 8;   // max MLP outputs (serial: E + 2)
constexpr int kMaxHH = 3;    // max number of HxH hidden layers (n_hidden - 1)
constexpr int kMaxIn = 32;   // max MLP inputs
constexpr int kMaxE = kMaxOut - 2;

struct Cfg {
  int arch;
  int B, T, t_stride, period_shift, ignore;
  int L, Lw, Le, E, W;
  int IN, IN4, XS;  // inputs, inputs padded to a multiple of 4 (or of 4*KQ0 in backward), smem row stride (IN4 + 4)
  int OUT;
  int NHH;          // number of HxH layers
  int w[kMaxHH + 3];// true widths: w[0]=IN, w[1..NHH+1]=hidden, w[NHH+2]=OUT
  int hidden_act;
  int lost, profit, has_edge, discrete, demand_layout, detach_input;
  float wub;
  // global (state_dict order) parameter offsets of each linear layer
  int gw[kMaxHH + 2], gb[kMaxHH + 2];
  int P;            // parameter count
  // shared-memory weight block offsets (floats)
  int s_wt0, s_b0, s_wth[kMaxHH], s_bh[kMaxHH], s_wo, s_bo, s_total;
  int tape_stride;  // floats per (t, scenario) row in the tape
  int ckpt;         // checkpoint interval K: the tape holds the state of periods 0, K, 2K, ... (K = 1: every period)
  float* ring;      // adjoint, K > 1: per-warp scratch [warp][K][32][tape_stride] for the recomputed states of a segment
  // tensor-core adjoint (precision != fp32): fp32 copies W[n][k] (row stride HS) of the HxH layers for the mma.sync
  // fragments, staged by the adjoint kernel only (after the s_total block the forward kernel uses)
  int tc, s_wn[kMaxHH], s_total_bwd;
  int K0, XSb, s_w0n;  // tc: first-layer contraction length padded to 8, row stride of the ADJOINT kernel's state rows
                       // (K0 + 4; = XS otherwise), fp32 copy W0[n][k] with row stride XSb
};

// lane = hidden unit / one scenario per warp form for training-size batches (rollout_small_unit.cu)
bool use_unit(const Cfg& c);
void set_unit_max_batch(int max_b);
void set_unit_group(int g);
int forward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const HdpoState* init,
                 float* cost_b, float* report_b, float* reward_tb, float* tape, const HdpoState& fin, void* stream);
int backward_unit(const Cfg& c, const float* params, const float* demands, const HdpoStatics* st, const float* tape,
                  float g_total, float g_report, float* partials, int p_stride, int* n_rows, void* stream);

bool supported(const HdpoRolloutDesc* d);
int build_cfg(const HdpoRolloutDesc* d, int in_pad_quantum, Cfg* c);
size_t workspace_bytes(const HdpoRolloutDesc* d);

int forward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
            const HdpoState* init, float* cost_b, float* report_b, float* reward_tb, double* totals,
            HdpoState* final_state, void* workspace, size_t workspace_bytes, void* stream);
int backward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st, float g_total,
             float g_report, float* grad_params, void* workspace, size_t workspace_bytes, void* stream);

}  // namespace small
}  // namespace hdpo
