// rollout_small.cu - K1/K2 for small policy nets (VanillaOneStore, VanillaSerial; hidden width <= 32).
//
// Replaces, for a whole batch and all T periods in ONE launch each:
//   forward : trainer.py:181-216 (simulate_batch loop) = neural_networks.py:200-214 / 319-355 + environment.py:110-299
//   backward: autograd of the above (trainer.py:173)
//
// Design (B200): warp-centric; fp32 FFMA throughout in parity mode (true fp32 like the reference's sgemm), and in the
// default tf32x3 mode the adjoint's 32-wide layers (recompute, dgrad, weight gradients) on warp-level tensor cores
// (mma.sync 3xTF32, mma32.cuh) - see the TC template parameter of small_bwd_kernel.
//   * one lane = one scenario (NS scenarios per lane in the forward); a warp owns tiles of 32*NS scenarios and
//     loops over tiles persistently; the four warps of a CTA only share the read-only weight block in shared
//     memory, so the period loop needs __syncwarp() only - never __syncthreads().
//   * per-scenario vectors (state, activations, adjoints) are ROWS in shared memory, stride 4*odd floats, so
//     a lane's float4 row accesses are conflict-free and other lanes can read rows for the cooperative
//     weight-gradient tiles.
//   * weights are staged once per CTA, transposed to [k][n] so that an LDS.128 broadcast feeds 4 FFMAs
//     (x NS scenarios) of 32 independent accumulators.
//   * the forward writes the state tape [T][B][IN4] (coalesced rows); the adjoint re-reads it in reverse,
//     recomputes the activations (no activation tape), back-propagates per scenario, and accumulates the
//     parameter gradient in REGISTERS as per-lane 4x8 tiles of dW over the warp's 32 scenarios; partial
//     gradients go to a per-warp slab and a second kernel reduces the slabs in fixed order (deterministic).
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace small {

// ------------------------------------------------------------------------------------------------------------
// host-side configuration
// ------------------------------------------------------------------------------------------------------------

static int pad_in(int in) {
  if (in <= 4) return 4;
  if (in <= 8) return 8;
  if (in <= 16) return 16;
  if (in <= 20) return 20;
  return 32;
}

bool supported(const HdpoRolloutDesc* d) {
  if (d->arch != HDPO_ARCH_VANILLA_ONE_STORE && d->arch != HDPO_ARCH_VANILLA_SERIAL) return false;
  // precision: fp32 = FFMA everywhere; tf32x3 / tf32 = the adjoint's 32-wide layers on mma.sync 3xTF32 (build_cfg: tc)
  const HdpoProblem& pb = d->pb;
  if (pb.S != 1) return false;
  const HdpoMlp& m = d->master;
  if (m.n_layers < 2 || m.n_layers > kMaxHH + 2) return false;
  for (int i = 1; i < m.n_layers; ++i)
    if (m.widths[i] > H || m.widths[i] < 1) return false;
  int in = pb.L + pb.W * pb.Lw + pb.E * pb.Le;
  if (m.widths[0] != in || in > kMaxIn) return false;
  if (d->arch == HDPO_ARCH_VANILLA_ONE_STORE) {
    if (pb.W != 0 || pb.E != 0 || m.widths[m.n_layers] != 1) return false;
  } else {
    if (pb.W != 1 || pb.E < 1 || pb.E > kMaxE || m.widths[m.n_layers] != pb.E + 2) return false;
  }
  if (m.out_act != HDPO_ACT_NONE) return false;
  return true;
}

int build_cfg(const HdpoRolloutDesc* d, int, Cfg* c) {
  const HdpoProblem& pb = d->pb;
  const HdpoMlp& m = d->master;
  c->arch = d->arch;
  c->B = pb.B;
  c->T = d->T;
  c->t_stride = d->t_stride;
  c->period_shift = d->period_shift;
  c->ignore = d->ignore_periods;
  c->L = pb.L;
  c->Lw = pb.W ? pb.Lw : 0;
  c->Le = pb.E ? pb.Le : 0;
  c->E = pb.E;
  c->W = pb.W;
  c->IN = m.widths[0];
  c->IN4 = pad_in(c->IN);
  c->XS = ((c->IN4 / 4) & 1) ? c->IN4 : c->IN4 + 4;
  c->OUT = m.widths[m.n_layers];
  c->NHH = m.n_layers - 2;
  for (int i = 0; i <= m.n_layers; ++i) c->w[i] = m.widths[i];
  c->hidden_act = m.hidden_act;
  c->lost = pb.lost_demand;
  c->profit = pb.maximize_profit;
  c->has_edge = pb.has_edge_cost;
  c->discrete = d->discrete_allocation;
  c->demand_layout = d->demand_layout;
  c->detach_input = (d->arch == HDPO_ARCH_VANILLA_SERIAL);  // neural_networks.py:329 torch.tensor(x) detaches
  c->wub = d->warehouse_upper_bound;
  int off = 0;
  for (int i = 0; i < m.n_layers; ++i) {
    c->gw[i] = off;
    off += m.widths[i + 1] * m.widths[i];
    c->gb[i] = off;
    off += m.widths[i + 1];
  }
  c->P = off;
  int s = 0;
  c->s_wt0 = s;
  s += c->IN4 * H;
  c->s_b0 = s;
  s += H;
  for (int i = 0; i < c->NHH; ++i) {
    c->s_wth[i] = s;
    s += H * H;
    c->s_bh[i] = s;
    s += H;
  }
  c->s_wo = s;
  s += kMaxOut * H;
  c->s_bo = s;
  s += kMaxOut;
  c->s_total = s;
#ifdef HDPO_EMU
  c->tc = 0;
#else
  c->tc = d->precision != HDPO_PREC_FP32 && c->NHH > 0;
#endif
  s = (s + 3) & ~3;
  for (int i = 0; i < c->NHH; ++i) {
    c->s_wn[i] = s;
    if (c->tc) s += H * HS;
  }
  c->K0 = (c->IN4 + 7) & ~7;
  c->XSb = c->XS;
  c->s_w0n = s;
  if (c->tc) {
    c->XSb = c->K0 + 4;  // adjoint rows hold the zero-padded K0 inputs of the mma first layer; stride stays 4 * odd
    s += H * c->XSb;
  }
  c->s_total_bwd = c->tc ? s : c->s_total;
  c->tape_stride = c->IN4;
  c->ckpt = d->checkpoint_interval > 1 ? d->checkpoint_interval : 1;
  if (c->ckpt > c->T && c->T > 0) c->ckpt = c->T;
  c->ring = nullptr;
  return HDPO_OK;
}


static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
// workspace = [state tape: ceil(T / K) checkpoints][per-warp gradient slabs][K > 1: per-warp segment rings]
static size_t tape_bytes(const Cfg& c, int save) {
  const size_t slots = (static_cast<size_t>(c.T) + c.ckpt - 1) / c.ckpt;
  return save ? align256(slots * c.B * c.tape_stride * sizeof(float)) : 0;
}
static size_t partial_bytes(const Cfg& c) {
  return align256(static_cast<size_t>(kMaxPartialRows) * ((c.P + 3) & ~3) * sizeof(float));
}
static size_t ring_bytes(const Cfg& c, int save) {
  if (!save || c.ckpt <= 1) return 0;
  // one ring per adjoint warp: the launch has at most ceil(B / 32) rounded up to a CTA, capped at kMaxPartialRows
  size_t warps = static_cast<size_t>(ceil_div(c.B, 32)) + kWarpsPerCta;
  if (warps > kMaxPartialRows) warps = kMaxPartialRows;
  return align256(warps * c.ckpt * 32 * c.tape_stride * sizeof(float));
}

size_t workspace_bytes(const HdpoRolloutDesc* d) {
  Cfg c;
  build_cfg(d, 0, &c);
  return tape_bytes(c, d->save_for_backward) + partial_bytes(c) + ring_bytes(c, d->save_for_backward) + 256;
}

// ------------------------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------------------------

static int sm_count() {
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

constexpr int kFwdNS = 2;

// With few scenario tiles (training batches of a few thousand scenarios) one warp per CTA spreads the tiles over all
// SMs and shortens the latency-bound per-tile time; big batches use 4 warps per CTA to share the staged weights.
static int pick_warps_per_cta(int n_tiles) {
  const int sms = sm_count();
  if (n_tiles <= 4 * sms) return 1;
  if (n_tiles <= 8 * sms) return 2;
  return kWarpsPerCta;
}

int forward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
            const HdpoState* init, float* cost_b, float* report_b, float* reward_tb, double* totals,
            HdpoState* final_state, void* workspace, size_t ws_bytes, void* stream) {
  Cfg c;
  build_cfg(d, 0, &c);
  float* tape = nullptr;
  if (d->save_for_backward) {
    HDPO_REQUIRE(workspace != nullptr, "save_for_backward needs a workspace");
    if (ws_bytes < workspace_bytes(d)) {
      set_error("workspace too small: %zu < %zu", ws_bytes, workspace_bytes(d));
      return HDPO_E_WORKSPACE;
    }
    tape = static_cast<float*>(workspace);
  }
  HdpoState fin = {nullptr, nullptr, nullptr};
  if (final_state) fin = *final_state;
  if (use_unit(c)) {
    int rc = forward_unit(c, params, demands, st, init, cost_b, report_b, reward_tb, tape, fin, stream);
    if (rc) return rc;
    if (totals) {
      auto kt = totals_kernel;
      HDPO_LAUNCH(kt, 1, 1024, 0, stream, static_cast<const float*>(cost_b), static_cast<const float*>(report_b), c.B,
                  totals);
      HDPO_LAUNCH_OK();
    }
    return HDPO_OK;
  }
  // two scenarios per lane amortise the weight loads when there are plenty of tiles; with few scenarios (training
  // batches of a few thousand, the 32768-scenario evaluation sets) one per lane doubles the warps that share the work
  const int ns = ceil_div(c.B, 32 * kFwdNS) < 4 * sm_count() ? 1 : kFwdNS;
  const int rows = 32 * ns;
  const int n_tiles = ceil_div(c.B, rows);
  const int wpc = pick_warps_per_cta(n_tiles);
  const size_t smem_max = (static_cast<size_t>((c.s_total + 3) & ~3) + static_cast<size_t>(kWarpsPerCta) * rows * (c.XS + HS)) *
                          sizeof(float);
  const size_t smem = (static_cast<size_t>((c.s_total + 3) & ~3) + static_cast<size_t>(wpc) * rows * (c.XS + HS)) * sizeof(float);
  const int ctas_needed = ceil_div(n_tiles, wpc);
  const int max_ctas = sm_count() * 3;
  const int grid = ctas_needed < max_ctas ? ctas_needed : max_ctas;
#define HDPO_FWD_LAUNCH(ARCH, NS)                                                                                      \
  do {                                                                                                                \
    auto k = small_fwd_kernel<ARCH, NS>;                                                                              \
    HDPO_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max)));   \
    HDPO_LAUNCH(k, grid, wpc * 32, smem, stream, c, params, demands, *st, *init, cost_b, report_b, reward_tb, tape,   \
                fin);                                                                                                 \
  } while (0)
  if (c.arch == HDPO_ARCH_VANILLA_ONE_STORE) {
    if (ns == 1) HDPO_FWD_LAUNCH(HDPO_ARCH_VANILLA_ONE_STORE, 1);
    else HDPO_FWD_LAUNCH(HDPO_ARCH_VANILLA_ONE_STORE, kFwdNS);
  } else {
    if (ns == 1) HDPO_FWD_LAUNCH(HDPO_ARCH_VANILLA_SERIAL, 1);
    else HDPO_FWD_LAUNCH(HDPO_ARCH_VANILLA_SERIAL, kFwdNS);
  }
#undef HDPO_FWD_LAUNCH
  HDPO_LAUNCH_OK();
  if (totals) {
    auto kt = totals_kernel;
    HDPO_LAUNCH(kt, 1, 1024, 0, stream, static_cast<const float*>(cost_b), static_cast<const float*>(report_b), c.B,
                totals);
    HDPO_LAUNCH_OK();
  }
  return HDPO_OK;
}

int backward(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st, float g_total,
             float g_report, float* grad_params, void* workspace, size_t ws_bytes, void* stream) {
  Cfg c;
  build_cfg(d, 0, &c);
  HDPO_REQUIRE(workspace != nullptr && d->save_for_backward,
               "backward needs the workspace of a forward run with save_for_backward = 1");
  if (ws_bytes < workspace_bytes(d)) {
    set_error("workspace too small: %zu < %zu", ws_bytes, workspace_bytes(d));
    return HDPO_E_WORKSPACE;
  }
  const float* tape = static_cast<const float*>(workspace);
  float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + tape_bytes(c, 1));
  if (c.ckpt > 1) c.ring = reinterpret_cast<float*>(static_cast<char*>(workspace) + tape_bytes(c, 1) + partial_bytes(c));
  const int p_stride = (c.P + 3) & ~3;
  if (use_unit(c)) {
    int n_rows = 0;
    int rc = backward_unit(c, params, demands, st, tape, g_total, g_report, partials, p_stride, &n_rows, stream);
    if (rc) return rc;
    auto kr = reduce_partials_kernel;
    HDPO_LAUNCH(kr, ceil_div(c.P, 256), 256, 0, stream, static_cast<const float*>(partials), n_rows, p_stride, c.P,
                grad_params);
    HDPO_LAUNCH_OK();
    return HDPO_OK;
  }
  const int n_tiles = ceil_div(c.B, 32);
  const int wpc = pick_warps_per_cta(n_tiles);
  const int ctas_needed = ceil_div(n_tiles, wpc);
  int max_ctas = sm_count() * 2 * (kWarpsPerCta / wpc);
  if (max_ctas * wpc > kMaxPartialRows) max_ctas = kMaxPartialRows / wpc;
  const int grid = ctas_needed < max_ctas ? ctas_needed : max_ctas;
  int rc;
#define HDPO_BWD_CASE(KQ0)                                                                                          \
  case KQ0:                                                                                                         \
    rc = (c.arch == HDPO_ARCH_VANILLA_ONE_STORE)                                                                    \
             ? launch_bwd_nhh<HDPO_ARCH_VANILLA_ONE_STORE, KQ0>(c, params, demands, st, tape, g_total, g_report,     \
                                                                partials, p_stride, grid, wpc, stream)              \
             : launch_bwd_nhh<HDPO_ARCH_VANILLA_SERIAL, KQ0>(c, params, demands, st, tape, g_total, g_report,        \
                                                             partials, p_stride, grid, wpc, stream);                \
    break;
  switch (c.IN4 / 4) {
    HDPO_BWD_CASE(1)
    HDPO_BWD_CASE(2)
    HDPO_BWD_CASE(4)
    HDPO_BWD_CASE(5)
    HDPO_BWD_CASE(8)
    default: set_error("unsupported padded input width %d", c.IN4); return HDPO_E_INVALID;
  }
#undef HDPO_BWD_CASE
  if (rc) return rc;
  auto kr = reduce_partials_kernel;
  HDPO_LAUNCH(kr, ceil_div(c.P, 256), 256, 0, stream, static_cast<const float*>(partials), grid * wpc,
              p_stride, c.P, grad_params);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

}  // namespace small
}  // namespace hdpo
