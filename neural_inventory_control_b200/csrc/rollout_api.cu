// rollout_api.cu - C ABI entry points of the fused rollout (K1/K2) and the host-buffer training step.
// Dispatch: small nets (hidden width <= 32, one store)  -> rollout_small.cu  (warp-centric SIMT, fp32)
//           everything else (wide MLPs, many stores)    -> rollout_wide.cu   (tile GEMM pipeline)
#include "rollout_small.cuh"
#include "rollout_wide.cuh"

namespace hdpo {

static int64_t mlp_params(const HdpoMlp& m) {
  int64_t n = 0;
  for (int i = 0; i < m.n_layers; ++i) n += static_cast<int64_t>(m.widths[i + 1]) * m.widths[i] + m.widths[i + 1];
  return n;
}

static int validate_mlp(const HdpoMlp& m, const char* name) {
  HDPO_REQUIRE(m.n_layers >= 1 && m.n_layers <= HDPO_MAX_LAYERS, "%s: n_layers=%d out of range", name, m.n_layers);
  for (int i = 0; i <= m.n_layers; ++i) HDPO_REQUIRE(m.widths[i] >= 1, "%s: width[%d]=%d", name, i, m.widths[i]);
  HDPO_REQUIRE(m.hidden_act >= HDPO_ACT_NONE && m.hidden_act <= HDPO_ACT_SOFTPLUS, "%s: bad hidden_act", name);
  HDPO_REQUIRE(m.out_act >= HDPO_ACT_NONE && m.out_act <= HDPO_ACT_SOFTPLUS, "%s: bad out_act", name);
  return HDPO_OK;
}

static int validate_desc(const HdpoRolloutDesc* d) {
  HDPO_REQUIRE(d != nullptr, "null descriptor");
  int rc = validate_problem(&d->pb);
  if (rc) return rc;
  HDPO_REQUIRE(d->T >= 1, "T=%d", d->T);
  HDPO_REQUIRE(d->period_shift >= 0 && d->t_stride >= d->T + d->period_shift,
               "demand time extent %d < T + period_shift = %d", d->t_stride, d->T + d->period_shift);
  HDPO_REQUIRE(d->ignore_periods >= 0, "ignore_periods=%d", d->ignore_periods);
  HDPO_REQUIRE(d->demand_layout == HDPO_DEMAND_BST || d->demand_layout == HDPO_DEMAND_TSB, "bad demand_layout");
  HDPO_REQUIRE(d->arch >= HDPO_ARCH_VANILLA_ONE_STORE && d->arch <= HDPO_ARCH_SYMMETRY_AWARE, "bad arch %d", d->arch);
  rc = validate_mlp(d->master, "master");
  if (rc) return rc;
  if (d->arch == HDPO_ARCH_SYMMETRY_AWARE) {
    rc = validate_mlp(d->store_net, "store");
    if (rc) return rc;
    rc = validate_mlp(d->warehouse_net, "warehouse");
    if (rc) return rc;
  }
  return HDPO_OK;
}

static int check_statics(const HdpoRolloutDesc* d, const HdpoStatics* st) {
  HDPO_REQUIRE(st && st->holding_costs && st->underage_costs && st->lead_times, "store statics missing");
  if (d->pb.W > 0) {
    HDPO_REQUIRE(st->warehouse_lead_times && st->warehouse_holding_costs, "warehouse statics missing");
    HDPO_REQUIRE(!d->pb.has_edge_cost || st->warehouse_edge_costs, "warehouse_edge_costs missing");
  }
  if (d->pb.E > 0) HDPO_REQUIRE(st->echelon_lead_times && st->echelon_holding_costs, "echelon statics missing");
  // the SymmetryAware store net reads the per-store demand mean / std as input features
  HDPO_REQUIRE(d->arch != HDPO_ARCH_SYMMETRY_AWARE || (st->mean && st->std),
               "symmetry_aware needs the 'mean' and 'std' static features (observation_params.include_static_features)");
  return HDPO_OK;
}

}  // namespace hdpo

using namespace hdpo;

extern "C" int64_t hdpo_param_count(const HdpoRolloutDesc* d) {
  if (!d) return -1;
  int64_t n = mlp_params(d->master);
  if (d->arch == HDPO_ARCH_SYMMETRY_AWARE) n += mlp_params(d->store_net) + mlp_params(d->warehouse_net);
  return n;
}

// ---- K4 in the path: demand generated on the device into the tail of the workspace (HdpoRolloutDesc.demand_source) ----
static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
static size_t generated_demand_bytes(const HdpoRolloutDesc* d) {
  if (d->demand_source == HDPO_DEMAND_FROM_ARGUMENT) return 0;
  return align256(static_cast<size_t>(d->t_stride) * d->pb.S * d->pb.B * sizeof(float));
}
static size_t inner_workspace_bytes(const HdpoRolloutDesc* d) {
  if (small::supported(d)) return small::workspace_bytes(d);
  if (wide::supported(d)) return wide::workspace_bytes(d);
  return 0;
}
// descriptor the kernels see (a plain time-major demand tensor) + where that tensor lives
static int resolve_demand(const HdpoRolloutDesc* d, const float** demands, void* workspace, size_t* workspace_bytes,
                          bool generate, void* stream, HdpoRolloutDesc* out) {
  *out = *d;
  if (d->demand_source == HDPO_DEMAND_FROM_ARGUMENT) return HDPO_OK;
  HDPO_REQUIRE(d->demand_source == HDPO_DEMAND_PHILOX_NORMAL || d->demand_source == HDPO_DEMAND_PHILOX_POISSON,
               "bad demand_source %d", d->demand_source);
  HDPO_REQUIRE(d->demand_mean && (d->demand_source == HDPO_DEMAND_PHILOX_POISSON || d->demand_std),
               "demand_source = Philox needs demand_mean (and demand_std for the normal sampler)");
  const size_t gen = generated_demand_bytes(d);
  HDPO_REQUIRE(workspace != nullptr, "the generated demand trace lives in the workspace");
  if (*workspace_bytes < gen + inner_workspace_bytes(d)) {
    set_error("workspace too small: %zu < %zu", *workspace_bytes, gen + inner_workspace_bytes(d));
    return HDPO_E_WORKSPACE;
  }
  *workspace_bytes -= gen;
  float* buf = reinterpret_cast<float*>(static_cast<char*>(workspace) + *workspace_bytes);
  out->demand_source = HDPO_DEMAND_FROM_ARGUMENT;
  out->demand_layout = HDPO_DEMAND_TSB;
  *demands = buf;
  if (!generate || d->pb.B == 0) return HDPO_OK;
  if (d->demand_source == HDPO_DEMAND_PHILOX_NORMAL)
    return hdpo_philox_normal(buf, d->pb.B, d->pb.S, d->t_stride, HDPO_DEMAND_TSB, d->demand_mean, d->demand_std, d->demand_rho,
                              d->demand_clip_at_zero, d->philox_seed, d->philox_offset, stream);
  return hdpo_philox_poisson(buf, d->pb.B, d->pb.S, d->t_stride, HDPO_DEMAND_TSB, d->demand_mean, d->philox_seed,
                             d->philox_offset, stream);
}

extern "C" size_t hdpo_rollout_workspace_bytes(const HdpoRolloutDesc* d) {
  if (validate_desc(d)) return 0;
  const size_t inner = inner_workspace_bytes(d);
  if (inner == 0) return 0;
  return align256(inner) + generated_demand_bytes(d);
}

extern "C" int hdpo_rollout_fwd(const HdpoRolloutDesc* d, const float* params, const float* demands,
                                const HdpoStatics* st, const HdpoState* init, float* cost_b, float* report_b,
                                float* reward_tb, double* totals, HdpoState* final_state, void* workspace,
                                size_t workspace_bytes, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  rc = check_statics(d, st);
  if (rc) return rc;
  HdpoRolloutDesc resolved;
  if ((rc = resolve_demand(d, &demands, workspace, &workspace_bytes, true, stream, &resolved))) return rc;
  d = &resolved;
  HDPO_REQUIRE(params && demands && init && init->store && cost_b, "null argument");
  HDPO_REQUIRE(d->pb.W == 0 || init->warehouse, "initial warehouse inventories missing");
  HDPO_REQUIRE(d->pb.E == 0 || init->echelon, "initial echelon inventories missing");
  if (d->pb.B == 0) return HDPO_OK;
  if (small::supported(d))
    return small::forward(d, params, demands, st, init, cost_b, report_b, reward_tb, totals, final_state, workspace,
                          workspace_bytes, stream);
  if (wide::supported(d))
    return wide::forward(d, params, demands, st, init, cost_b, report_b, reward_tb, totals, final_state, workspace,
                         workspace_bytes, stream);
  set_error("no fused rollout for this architecture / shape (arch=%d); use the per-step path", d->arch);
  return HDPO_E_INVALID;
}

extern "C" int hdpo_rollout_bwd(const HdpoRolloutDesc* d, const float* params, const float* demands,
                                const HdpoStatics* st, float g_total, float g_report, float* grad_params,
                                void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  rc = check_statics(d, st);
  if (rc) return rc;
  HdpoRolloutDesc resolved;  // demand generated by the forward call: read it back from the workspace
  if ((rc = resolve_demand(d, &demands, workspace, &workspace_bytes, false, stream, &resolved))) return rc;
  d = &resolved;
  HDPO_REQUIRE(params && demands && grad_params, "null argument");
  HDPO_REQUIRE(!d->discrete_allocation, "discrete_allocation is forward-only (trainer.py:201-202 rounds under no_grad)");
  if (d->pb.B == 0) {
    HDPO_CUDA_OK(cudaMemsetAsync(grad_params, 0, sizeof(float) * hdpo_param_count(d), static_cast<cudaStream_t>(stream)));
    return HDPO_OK;
  }
  if (small::supported(d))
    return small::backward(d, params, demands, st, g_total, g_report, grad_params, workspace, workspace_bytes, stream);
  if (wide::supported(d))
    return wide::backward(d, params, demands, st, g_total, g_report, grad_params, workspace, workspace_bytes, stream);
  set_error("no fused rollout for this architecture / shape (arch=%d)", d->arch);
  return HDPO_E_INVALID;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer training step: H2D of everything the batch needs, forward + adjoint, D2H of the results
// ---------------------------------------------------------------------------------------------------------
namespace {
struct HostLayout {
  size_t params, grad, demands, hold, under, lead, wlead, whold, wedge, elead, ehold, mean, std_, store, wh, ech, adj,
      cost, report, totals, inner, total;
};
size_t a256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
HostLayout host_layout(const HdpoRolloutDesc* d) {
  const HdpoProblem& pb = d->pb;
  const size_t B = pb.B, S = pb.S, W = pb.W, E = pb.E, Wc = W > 0 ? W : 1, f = sizeof(float);
  const size_t P = static_cast<size_t>(hdpo_param_count(d));
  HostLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += a256(bytes);
    return at;
  };
  l.params = take(P * f);
  l.grad = take(P * f);
  l.demands = take(d->demand_source == HDPO_DEMAND_FROM_ARGUMENT ? B * S * d->t_stride * f : 2 * S * f);  // or [mean | std]
  l.hold = take(B * S * f);
  l.under = take(B * S * f);
  l.lead = take(B * S * Wc * f);
  l.wlead = take(B * W * f);
  l.whold = take(B * W * f);
  l.wedge = take(B * W * f);
  l.elead = take(B * E * f);
  l.ehold = take(B * E * f);
  l.mean = take(B * S * f);
  l.std_ = take(B * S * f);
  l.store = take(B * S * pb.L * f);
  l.wh = take(B * W * pb.Lw * f);
  l.ech = take(B * E * pb.Le * f);
  l.adj = take(W * S * sizeof(int32_t));
  l.cost = take(B * f);
  l.report = take(B * f);
  l.totals = take(2 * sizeof(double));
  l.inner = o;
  l.total = o + a256(hdpo_rollout_workspace_bytes(d));
  return l;
}
}  // namespace

extern "C" size_t hdpo_rollout_host_workspace_bytes(const HdpoRolloutDesc* d) {
  if (validate_desc(d)) return 0;
  return host_layout(d).total;
}

extern "C" int hdpo_rollout_train_host(const HdpoRolloutDesc* d, const float* h_params, const float* h_demands,
                                       const HdpoStatics* h_st, const HdpoState* h_init, const int32_t* h_adjacency,
                                       double* h_totals, float* h_grad_params, void* d_workspace,
                                       size_t workspace_bytes, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  rc = check_statics(d, h_st);
  if (rc) return rc;
  const bool gen = d->demand_source != HDPO_DEMAND_FROM_ARGUMENT;
  HDPO_REQUIRE(h_params && (h_demands || gen) && h_init && h_init->store && h_totals && h_grad_params && d_workspace,
               "null argument");
  HDPO_REQUIRE(!gen || d->demand_mean, "demand_source = Philox needs demand_mean (host pointer for the host step)");
  const HostLayout l = host_layout(d);
  if (workspace_bytes < l.total) {
    set_error("host-step workspace too small: %zu < %zu", workspace_bytes, l.total);
    return HDPO_E_WORKSPACE;
  }
  const HdpoProblem& pb = d->pb;
  const size_t B = pb.B, S = pb.S, W = pb.W, E = pb.E, Wc = W > 0 ? W : 1, f = sizeof(float);
  const size_t P = static_cast<size_t>(hdpo_param_count(d));
  char* base = static_cast<char*>(d_workspace);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
    if (!src || bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(base + off, src, bytes, cudaMemcpyHostToDevice, s);
  };
  HDPO_CUDA_OK(up(l.params, h_params, P * f));
  if (!gen) {
    HDPO_CUDA_OK(up(l.demands, h_demands, B * S * d->t_stride * f));
  } else {  // no demand crosses the bus: S means (+ S standard deviations) parameterise the on-device sampler
    HDPO_CUDA_OK(up(l.demands, d->demand_mean, S * f));
    HDPO_CUDA_OK(up(l.demands + S * f, d->demand_std, S * f));
  }
  HDPO_CUDA_OK(up(l.hold, h_st->holding_costs, B * S * f));
  HDPO_CUDA_OK(up(l.under, h_st->underage_costs, B * S * f));
  HDPO_CUDA_OK(up(l.lead, h_st->lead_times, B * S * Wc * f));
  HDPO_CUDA_OK(up(l.wlead, h_st->warehouse_lead_times, B * W * f));
  HDPO_CUDA_OK(up(l.whold, h_st->warehouse_holding_costs, B * W * f));
  HDPO_CUDA_OK(up(l.wedge, h_st->warehouse_edge_costs, B * W * f));
  HDPO_CUDA_OK(up(l.elead, h_st->echelon_lead_times, B * E * f));
  HDPO_CUDA_OK(up(l.ehold, h_st->echelon_holding_costs, B * E * f));
  HDPO_CUDA_OK(up(l.mean, h_st->mean, B * S * f));
  HDPO_CUDA_OK(up(l.std_, h_st->std, B * S * f));
  HDPO_CUDA_OK(up(l.store, h_init->store, B * S * pb.L * f));
  HDPO_CUDA_OK(up(l.wh, h_init->warehouse, B * W * pb.Lw * f));
  HDPO_CUDA_OK(up(l.ech, h_init->echelon, B * E * pb.Le * f));
  HDPO_CUDA_OK(up(l.adj, h_adjacency, W * S * sizeof(int32_t)));
  auto dp = [&](size_t off, const void* host) -> float* { return host ? reinterpret_cast<float*>(base + off) : nullptr; };
  HdpoStatics st = {dp(l.hold, h_st->holding_costs), dp(l.under, h_st->underage_costs), dp(l.lead, h_st->lead_times),
                    dp(l.wlead, h_st->warehouse_lead_times), dp(l.whold, h_st->warehouse_holding_costs),
                    dp(l.wedge, h_st->warehouse_edge_costs), dp(l.elead, h_st->echelon_lead_times),
                    dp(l.ehold, h_st->echelon_holding_costs), dp(l.mean, h_st->mean), dp(l.std_, h_st->std)};
  HdpoState init = {dp(l.store, h_init->store), dp(l.wh, h_init->warehouse), dp(l.ech, h_init->echelon)};
  HdpoRolloutDesc dd = *d;
  dd.save_for_backward = 1;
  dd.adjacency = h_adjacency ? reinterpret_cast<const int32_t*>(base + l.adj) : nullptr;
  if (gen) {
    dd.demand_mean = reinterpret_cast<const float*>(base + l.demands);
    dd.demand_std = d->demand_std ? reinterpret_cast<const float*>(base + l.demands) + S : nullptr;
  }
  float* params = reinterpret_cast<float*>(base + l.params);
  float* grad = reinterpret_cast<float*>(base + l.grad);
  float* demands = reinterpret_cast<float*>(base + l.demands);
  double* totals = reinterpret_cast<double*>(base + l.totals);
  void* inner = base + l.inner;
  const size_t inner_bytes = l.total - l.inner;
  rc = hdpo_rollout_fwd(&dd, params, demands, &st, &init, reinterpret_cast<float*>(base + l.cost),
                        reinterpret_cast<float*>(base + l.report), nullptr, totals, nullptr, inner, inner_bytes, stream);
  if (rc) return rc;
  const float g_total = 1.0f / (static_cast<float>(B) * static_cast<float>(d->T) * static_cast<float>(S));  // trainer.py:169
  rc = hdpo_rollout_bwd(&dd, params, demands, &st, g_total, 0.f, grad, inner, inner_bytes, stream);
  if (rc) return rc;
  HDPO_CUDA_OK(cudaMemcpyAsync(h_totals, totals, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  HDPO_CUDA_OK(cudaMemcpyAsync(h_grad_params, grad, P * f, cudaMemcpyDeviceToHost, s));
  HDPO_CUDA_OK(cudaStreamSynchronize(s));
  return HDPO_OK;
}
