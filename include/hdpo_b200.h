/*
 * hdpo_b200.h - C ABI of the B200-native HDPO rollout engine (libhdpo_b200.so).
 *
 * The reference (MatiasAlvo/Neural_inventory_control) has no FFI layer: its hot path is Python calling
 * PyTorch eager ops.  This header is the boundary a maintainer would bind (ctypes stub in
 * INTEGRATION.md) to replace exactly that path:
 *
 *   hdpo_step_fwd / hdpo_step_bwd        <- environment.py:110-299,391-434  Simulator.step (+ its autograd)
 *   hdpo_allocation_shift                <- environment.py:77-101           initialize_shifts_for_allocation_put
 *   hdpo_rollout_fwd                     <- trainer.py:181-216              Trainer.simulate_batch
 *                                           neural_networks.py:195-214,314-427  Vanilla{OneStore,Serial,Warehouse}.forward
 *                                           neural_networks.py:111-166      feasibility projections
 *                                           loss_functions.py:10-11         PolicyLoss (reward.sum())
 *   hdpo_rollout_bwd                     <- trainer.py:169-173              mean_loss.backward()
 *   hdpo_rollout_train_host              <- trainer.py:155-173              one batch, host buffers in / host results out
 *   hdpo_philox_normal / _poisson        <- data_handling.py:178-211        Scenario.generate_*_demand (statistical parity)
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.  All arrays are dense, row-major, float32
 *     unless stated; "device" pointers are CUDA device memory of the current device.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) except the *_host entry point, which synchronises the stream before returning.
 *   - return value: 0 = ok, negative = error (HDPO_E_*); hdpo_last_error() returns a thread-local,
 *     human-readable message for the last failing call.  Nothing throws across the boundary.
 *   - the caller owns every buffer.  The only scratch memory is the explicit `workspace`, whose size
 *     is queried first with hdpo_rollout_workspace_bytes().  No hidden allocation, no CPU fallback.
 *   - thread-compatible: concurrent calls must use distinct streams AND distinct workspaces.
 */
#ifndef HDPO_B200_H_
#define HDPO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDPO_ABI_VERSION 2
#define HDPO_MAX_LAYERS 8 /* linear layers per MLP */

enum {
  HDPO_OK = 0,
  HDPO_E_INVALID = -1,     /* bad descriptor / null pointer / unsupported shape */
  HDPO_E_CUDA = -2,        /* a CUDA runtime call failed (message has the cudaError string) */
  HDPO_E_NO_DEVICE = -3,   /* no CUDA device / not an sm_100 device */
  HDPO_E_WORKSPACE = -4    /* workspace too small */
};

/* policy architectures with a fused rollout (neural_networks.py:1519-1536 registry names) */
enum {
  HDPO_ARCH_VANILLA_ONE_STORE = 0, /* neural_networks.py:195-214 */
  HDPO_ARCH_VANILLA_SERIAL = 1,    /* neural_networks.py:314-355 */
  HDPO_ARCH_VANILLA_WAREHOUSE = 2, /* neural_networks.py:358-427 */
  HDPO_ARCH_SYMMETRY_AWARE = 3     /* SURVEY.md 2.3 (recovered from stale bytecode) */
};

/* activation ids (neural_networks.py:36-43) */
enum { HDPO_ACT_NONE = 0, HDPO_ACT_ELU = 1, HDPO_ACT_RELU = 2, HDPO_ACT_TANH = 3, HDPO_ACT_SIGMOID = 4, HDPO_ACT_SOFTPLUS = 5 };

/* demand tensor layouts */
enum {
  HDPO_DEMAND_BST = 0, /* [B, S, t_stride]  the reference layout (data_handling.py:59, environment.py:171-177) */
  HDPO_DEMAND_TSB = 1  /* [t_stride, S, B]  time-major, what hdpo_philox_* writes; fully coalesced */
};

/* where the demand of a rollout comes from (HdpoRolloutDesc.demand_source) */
enum {
  HDPO_DEMAND_FROM_ARGUMENT = 0,  /* the `demands` argument (device / host pointer) */
  HDPO_DEMAND_PHILOX_NORMAL = 1,  /* generated on the device: hdpo_philox_normal(mean, std, rho, clip) */
  HDPO_DEMAND_PHILOX_POISSON = 2  /* generated on the device: hdpo_philox_poisson(mean) */
};

/* matmul precision of the policy MLP */
enum {
  HDPO_PREC_FP32 = 0,   /* FFMA, fp32 everywhere (parity mode) */
  HDPO_PREC_TF32X3 = 1, /* tensor cores with the 3-pass tf32 split (hi*hi + hi*lo + lo*hi), fp32 accumulate: tcgen05
                           tile GEMMs for the wide policy MLPs; warp-level mma.sync for the 32-wide nets inside the
                           adjoint kernels (small nets, SymmetryAware heads). fp32-grade: same parity bar as FP32 */
  HDPO_PREC_TF32 = 2    /* tcgen05 kind::tf32, single pass (throughput mode; NOT within the 1e-5 bar) */
};

/* ------------------------------------------------------------------------------------------------
 * Problem shape shared by the per-step and the rollout entry points.
 * (problem_params + tensor shapes of Scenario.get_data(), data_handling.py:54-81)
 * ---------------------------------------------------------------------------------------------- */
typedef struct HdpoProblem {
  int32_t B;               /* scenarios in the batch */
  int32_t S;               /* n_stores >= 1 */
  int32_t W;               /* n_warehouses (0 allowed); action width Wc = max(W,1) */
  int32_t E;               /* n_extra_echelons */
  int32_t L, Lw, Le;       /* pipeline lengths: store_inventories.shape[2], warehouse..., echelon... (>= 2) */
  int32_t lost_demand;     /* problem_params['lost_demand'] */
  int32_t maximize_profit; /* problem_params['maximize_profit'] */
  int32_t has_edge_cost;   /* 'warehouse_edge_costs' present (environment.py:254) */
} HdpoProblem;

/* Per-scenario constants, all device pointers, reference layouts (float32; lead times float-encoded
 * integers exactly as Scenario.get_data() emits them, data_handling.py:81). Unused ones may be NULL. */
typedef struct HdpoStatics {
  const float* holding_costs;           /* [B,S] */
  const float* underage_costs;          /* [B,S] */
  const float* lead_times;              /* [B,S,Wc] */
  const float* warehouse_lead_times;    /* [B,W] */
  const float* warehouse_holding_costs; /* [B,W] */
  const float* warehouse_edge_costs;    /* [B,W] or NULL */
  const float* echelon_lead_times;      /* [B,E] */
  const float* echelon_holding_costs;   /* [B,E] */
  const float* mean;                    /* [B,S] symmetry-aware only */
  const float* std;                     /* [B,S] symmetry-aware only */
} HdpoStatics;

/* Inventory state, reference layouts. */
typedef struct HdpoState {
  float* store;     /* [B,S,L]  */
  float* warehouse; /* [B,W,Lw] or NULL */
  float* echelon;   /* [B,E,Le] or NULL */
} HdpoState;

typedef struct HdpoAction {
  float* stores;     /* [B,S,Wc] */
  float* warehouses; /* [B,W,1] or NULL */
  float* echelons;   /* [B,E,1] or NULL */
} HdpoAction;

/* ------------------------------------------------------------------------------------------------
 * K3 - one simulator period for an arbitrary policy (environment.py:110-169).
 * ---------------------------------------------------------------------------------------------- */

/* demand[b*demand_stride_b + s*demand_stride_s] is the current demand of (b,s): pass the base pointer
 * already offset to column t+period_shift; strides are in elements.
 * Writes the next state into `next` (must not alias `cur`) and reward[B]. A NON-ZERO order whose lead time is outside
 * [1, pipeline length] lands where the reference's flat put (environment.py:422-432) puts it: in a neighbouring node's
 * pipeline (lead 0 = last slot of the previous node); orders that would fall outside the tensor are dropped. */
int hdpo_step_fwd(const HdpoProblem* pb, const HdpoStatics* st, const HdpoState* cur, const HdpoAction* act,
                  const float* demand, int64_t demand_stride_b, int64_t demand_stride_s, HdpoState* next,
                  float* reward, void* stream);

/* Adjoint of hdpo_step_fwd.  g_next: adjoint wrt `next` (NULL members = zero); g_reward[B];
 * outputs g_cur (adjoint wrt `cur`) and g_act (adjoint wrt the action).  `cur`, `act`, `demand` are the
 * forward inputs.  Sub-gradient conventions follow torch autograd (see DESIGN.md "kinks"). */
int hdpo_step_bwd(const HdpoProblem* pb, const HdpoStatics* st, const HdpoState* cur, const HdpoAction* act,
                  const float* demand, int64_t demand_stride_b, int64_t demand_stride_s, const HdpoState* g_next,
                  const float* g_reward, HdpoState* g_cur, HdpoAction* g_act, void* stream);

/* shift[b,n] = b*(len*n_nodes) + n*len as int64, bit-exact with environment.py:77-101. */
int hdpo_allocation_shift(int64_t* shift, int32_t B, int32_t n_nodes, int32_t len, void* stream);

/* Batch assembly on the device: dst[i][0..row_floats) = src[idx[i]][0..row_floats) for i < n_rows (idx int64).
 * Replaces the DataLoader's per-sample collate + per-batch H2D copy (data_handling.py:385-395, trainer.py:155-156)
 * once the dataset is resident in HBM. */
int hdpo_gather_rows(float* dst, const float* src, const int64_t* idx, int64_t n_rows, int64_t row_floats,
                     void* stream);

/* One fused Adam step over a flat parameter vector (replaces optimizer.step() of trainer.py:177 for torch.optim.Adam
 * without amsgrad; same arithmetic and operation order as torch/optim/adam.py::_single_tensor_adam). `step` is the
 * 1-based step count AFTER this update; exp_avg / exp_avg_sq are the optimizer's moment vectors (updated in place). */
int hdpo_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int64_t step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1 / K2 - fused T-period rollout and its reverse-time adjoint.
 * ---------------------------------------------------------------------------------------------- */
typedef struct HdpoMlp {
  int32_t n_layers;                     /* number of Linear layers (>= 1) */
  int32_t widths[HDPO_MAX_LAYERS + 1];  /* widths[0] = input size, widths[i+1] = out features of layer i */
  int32_t hidden_act;                   /* HDPO_ACT_* after every layer but the last */
  int32_t out_act;                      /* HDPO_ACT_* after the last layer */
} HdpoMlp;

typedef struct HdpoRolloutDesc {
  HdpoProblem pb;
  int32_t arch;               /* HDPO_ARCH_* */
  int32_t T;                  /* periods to simulate */
  int32_t t_stride;           /* time extent of the demand tensor (>= T + period_shift) */
  int32_t period_shift;       /* observation_params['demand']['period_shift'] */
  int32_t ignore_periods;     /* periods excluded from the reported loss (trainer.py:209) */
  int32_t demand_layout;      /* HDPO_DEMAND_* */
  int32_t discrete_allocation;/* round actions half-to-even (trainer.py:201-202); forward only */
  int32_t transshipment;      /* nn_params['transshipment'] (no hold logit / no clip at 1) */
  int32_t precision;          /* HDPO_PREC_* */
  int32_t save_for_backward;  /* 1: forward writes the state tape into the workspace */
  float warehouse_upper_bound;/* neural_networks.py:1538-1546 */
  float prop_eps;             /* symmetry-aware proportional allocation epsilon (1e-15 old / 1e-10 current) */
  HdpoMlp master;             /* vanilla_*: the 'master' net; symmetry_aware: the 'context' net */
  HdpoMlp store_net;          /* symmetry_aware only */
  HdpoMlp warehouse_net;      /* symmetry_aware only */
  const int32_t* adjacency;   /* device [W,S] 0/1, NULL = fully connected (neural_networks.py:383-390) */
  /* K4 wired into the path (ABI 2): with demand_source != 0 the `demands` argument may be NULL - the [t_stride, S, B]
   * trace is generated on the device by the Philox sampler (element (t, s, b) <-> counter philox_offset + index / 4,
   * exactly what hdpo_philox_normal / _poisson write for layout HDPO_DEMAND_TSB) into the workspace, once per forward
   * call, and read again by the adjoint. Replaces data_handling.py:178-211 + the per-batch H2D copy (trainer.py:156)
   * for synthetic settings. demand_mean / demand_std: [S] (device pointers; HOST pointers for *_train_host). */
  int32_t demand_source;      /* HDPO_DEMAND_* source */
  int32_t demand_clip_at_zero;
  float demand_rho;           /* one-factor correlation of the normal sampler */
  int32_t checkpoint_interval;/* K > 1: the forward keeps the state of every K-th period only and the adjoint re-runs
                                 the K - 1 periods in between (recomputation checkpointing: tape bytes / K for one
                                 extra forward pass; small-net path). 0 / 1: the state of every period is taped */
  uint64_t philox_seed, philox_offset;
  const float* demand_mean;
  const float* demand_std;
} HdpoRolloutDesc;

/* Number of float parameters the descriptor's nets hold, in state_dict order:
 * for each net (master|context, store, warehouse), for each layer: weight [out,in] then bias [out]. */
int64_t hdpo_param_count(const HdpoRolloutDesc* d);

/* Bytes of workspace hdpo_rollout_fwd (with save_for_backward) + hdpo_rollout_bwd need. */
size_t hdpo_rollout_workspace_bytes(const HdpoRolloutDesc* d);

/* Forward rollout.  params: device flat parameter vector (hdpo_param_count floats).
 * demands: device, layout per d->demand_layout.  init: initial inventories (read only).
 * Outputs (device): cost_b[B] = sum_t reward[t,b]; report_b[B] = same for t >= ignore_periods (may be NULL);
 * reward_tb [T,B] optional (NULL to skip); totals[2] (double) = {sum_b cost_b, sum_b report_b};
 * final: state after T periods (members may be NULL). */
int hdpo_rollout_fwd(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
                     const HdpoState* init, float* cost_b, float* report_b, float* reward_tb, double* totals,
                     HdpoState* final_state, void* workspace, size_t workspace_bytes, void* stream);

/* Reverse-time adjoint of the forward that last filled `workspace` (same d, params, demands, st).
 * dLoss/dreward[t,b] = g_total + (t >= ignore_periods ? g_report : 0).
 * grad_params (device, hdpo_param_count floats) is OVERWRITTEN with dLoss/dparams. */
int hdpo_rollout_bwd(const HdpoRolloutDesc* d, const float* params, const float* demands, const HdpoStatics* st,
                     float g_total, float g_report, float* grad_params, void* workspace, size_t workspace_bytes,
                     void* stream);

/* One training batch end to end with HOST buffers (pinned recommended): copies params, demands, statics
 * and initial state host->device into the workspace, runs forward + adjoint with
 * g_total = 1/(B*T*S) (trainer.py:169), copies totals[2] and grad_params back, synchronises.
 * Device scratch = hdpo_rollout_host_workspace_bytes(d). */
size_t hdpo_rollout_host_workspace_bytes(const HdpoRolloutDesc* d);
int hdpo_rollout_train_host(const HdpoRolloutDesc* d, const float* h_params, const float* h_demands,
                            const HdpoStatics* h_st, const HdpoState* h_init, const int32_t* h_adjacency,
                            double* h_totals, float* h_grad_params, void* d_workspace, size_t workspace_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4 - on-device Philox4x32-10 demand sampler (counter-based: element i of a call draws from counter
 * (offset + i/4), key = seed).  Output layout HDPO_DEMAND_TSB [T,S,B] or BST via `layout`.
 * normal: mean[s] + std[s]*z, optional one-factor correlation rho (cov_ij = rho*std_i*std_j, i != j,
 * data_handling.py:193-201), optional clip at 0 (data_handling.py:145-146).
 * ---------------------------------------------------------------------------------------------- */
int hdpo_philox_normal(float* out, int32_t B, int32_t S, int32_t T, int32_t layout, const float* mean,
                       const float* std, float rho, int32_t clip_at_zero, uint64_t seed, uint64_t offset,
                       void* stream);
int hdpo_philox_poisson(float* out, int32_t B, int32_t S, int32_t T, int32_t layout, const float* mean,
                        uint64_t seed, uint64_t offset, void* stream);
/* The raw stream: out[4*g .. 4*g+3] = Philox4x32-10(counter = offset + g, key = seed), g < n_groups.
 * Bit-exact with the Random123 known-answer vectors. */
int hdpo_philox_raw(uint32_t* out, uint64_t n_groups, uint64_t seed, uint64_t offset, void* stream);

/* Test hook for the tcgen05 tile GEMM: C[M,N] = A[M,K] * B[N,K]^T (n_pass 3 = 3xTF32, 1 = TF32); dense row-major
 * device arrays, M % 128 == 0, N % 64 == 0, K % 32 == 0; scratch = 2*(M*K + N*K) floats of device memory. */
int hdpo_debug_gemm_tc(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t n_pass,
                       float* scratch, void* stream);
/* Same for the weight-gradient (MN-major) form: C[M,N] = A[K,M]^T * B[K,N], A / B row-major [K][M] / [K][N];
 * K % k_per_split == 0, k_per_split % 32 == 0; scratch = 2*(K*M + K*N) + (K/k_per_split)*M*N floats. */
int hdpo_debug_gemm_tc_wgrad(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                             int32_t k_per_split, int32_t n_pass, float* scratch, void* stream);

/* The forward hidden-layer GEMM exactly as the wide rollout launches it, for timing it alone (bench.py `roofline`) and
 * for per-CTA clock stamps (tools/gemm_timeline.py): C (+ C_lo) = epilogue(A[M,K] * B[N,K]^T) with A / B already split
 * in `scratch` by a previous hdpo_debug_gemm_tc call on the same operands; epi = 0: bias + ELU + (hi, lo) outputs
 * (C_lo and bias required), epi = 4: plain store. dbg_clock: device array of 8 * n_ctas int64 (clock64 stamps). */
int hdpo_debug_gemm_tc_timeline(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                                int32_t n_pass, float* scratch, long long* dbg_clock, void* stream, int32_t epi,
                                float* C_lo, const float* bias);
/* Device-side trace buffer for tools/trace_step.py (one record per CTA of the wide path); NULL disables it. */
int hdpo_debug_set_trace(unsigned long long* buf, int64_t capacity);
/* Per-role event trace of the persistent wide sweeps (tools/wp_trace.py): buf = 74 pairs * 4 roles * cap_per_role
 * records of {tag, globaltimer ns} (uint64 pairs); NULL disables it. */
int hdpo_debug_set_wp_trace(unsigned long long* buf, int32_t cap_per_role);
/* Opt-in persistent one-launch sweeps of the wide path (wide_persist.cu); default off (env HDPO_WIDE_PERSIST=1). The
 * workspace size depends on it: query hdpo_rollout_workspace_bytes after switching. */
int hdpo_debug_set_wide_persist(int32_t on);
/* Routing threshold of the multi-tile CTA-pair GEMM of the wide path: min_tiles > 0 = fewest 256-row tiles of a layer
 * launch that goes there (1 = every tensor-core GEMM), 0 = never, < 0 = default (off unless HDPO_TC_MULTI=1). */
int hdpo_debug_set_tc_multi(int32_t min_tiles);
/* Two-CTAs-per-SM tile forms of the per-period tensor-core GEMMs (4 x 64 TMEM columns and a 2-stage ring per CTA):
 * 0 = off (default), 1 = 256 x 64 CTA-pair tiles where the epilogue exists in that form, 2 = 128 x 64 single-CTA tiles,
 * < 0 = back to HDPO_TC_OCC2. Opt-in: slower than the 256 x 128 pair tiles at 8192 scenarios. */
int hdpo_debug_set_tc_occ2(int32_t mode);
/* Weight-gradient GEMMs of the wide path cut into groups of `group` periods and run on a low-priority stream while the
 * adjoint sweep is still going: mode 1 = on, 0 = off, < 0 = default (on when the batch is ONE chunk); group <= 0 keeps
 * the current group size (HDPO_WIDE_WG_GROUP, 5). Changes hdpo_rollout_workspace_bytes. */
int hdpo_debug_set_wide_wg_overlap(int32_t mode, int32_t group);
/* Split-K of the two thin GEMMs on the wide path's per-period chain (output layer: partial products summed by the forward
 * head; first-layer dgrad: partials added by the next adjoint head): 1 = on, 0 = off, < 0 = default (on when the batch is
 * ONE chunk, where the chain is latency-bound: 1024 scenarios 6.25 -> 5.71 ms per step). Changes the workspace size and the
 * summation order of those two layers. */
int hdpo_debug_set_wide_ksplit(int32_t mode);
/* Largest batch (scenarios) that the small-net rollout runs in its one-scenario-per-warp form (rollout_small_unit.cu);
 * larger batches use the 32-scenarios-per-warp form. > 0 sets it, 0 = never, < 0 = default (HDPO_SMALL_UNIT_MAX, else
 * 4096 one-store / 2048 serial). */
int hdpo_debug_set_small_unit(int32_t max_batch);
/* Scenarios per warp of that form: 1, 2 or 4 (weights loaded once per warp for all of them, their policy heads on
 * lanes 0..G-1); anything else = chosen by the batch size (HDPO_SMALL_UNIT_G). */
int hdpo_debug_set_small_unit_group(int32_t g);

/* misc */
const char* hdpo_last_error(void);
int hdpo_abi_version(void);
/* Launch statistics since process start: number of kernels this library has launched. */
int64_t hdpo_kernel_launch_count(void);
/* SM count / name of the current device (fails with HDPO_E_NO_DEVICE when none). */
int hdpo_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char* name, int32_t name_len);

#ifdef __cplusplus
}
#endif
#endif /* HDPO_B200_H_ */
