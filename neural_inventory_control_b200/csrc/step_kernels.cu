// step_kernels.cu - K3: one simulator period (and its adjoint) for ARBITRARY policies.
//
// Replaces the ~60 ATen launches + 2-4 host syncs of Simulator.step (environment.py:110-169) and of its
// autograd with one kernel each. Two thread mappings, both with order-deterministic per-scenario sums (reward over
// stores, warehouse draw-down over stores):
//   * networks with >= kStepWarpMinStores (4) stores: ONE WARP per scenario, lanes over the stores (coalesced rows of the
//     [B, S, L] / [B, S, W] tensors, fixed-shape shuffle trees for the sums), warehouse w on lane w, echelons on lane 0;
//   * one-store / serial settings: one thread per scenario walking its few nodes sequentially.
// The fused rollout kernels (rollout_small.cu / rollout_wide.cu) are the fast path; this is the generic one that keeps
// `Simulator.step` a drop-in for any torch policy (GNN, DataDrivenNet, quantile and base-stock policies).
#include "hdpo_internal.cuh"

#include <cstdlib>

namespace hdpo {

// pipeline update of environment.py:391-434 for one node: dst[0] = post + src[1]; dst[k] = src[k+1]; dst[len-1] = 0
__device__ __forceinline__ void shift_pipeline(const float* __restrict__ src, float* __restrict__ dst, int len,
                                               float post) {
  dst[0] = post + src[1];
  for (int k = 1; k < len - 1; ++k) dst[k] = src[k + 1];
  dst[len - 1] = 0.f;
}

// the masked put(accumulate=True) of environment.py:422-432: exact zeros are skipped; slot = lead - 1.
// The reference's put works on the FLATTENED [B, nodes, len] tensor at index shift + lead - 1, so a non-zero order whose
// lead time is outside [1, len] lands in a NEIGHBOURING node's pipeline (lead 0: the last slot of the previous node;
// flat index -1 wraps to the tensor's last element). The shipped GNN policy does this on many_warehouses_lost_demand
// (its ragged edge -> column mapping pairs allocations with the lead-time-0 entries of unconnected pairs), so the
// behaviour is reproduced: in-range orders land here, out-of-range ones in stray_orders_kernel after this kernel.
__device__ __forceinline__ void land_order(float* __restrict__ dst, int len, float amount, float lead, int* stray) {
  if (amount != 0.f) {
    const int slot = static_cast<int>(lead) - 1;
    if (slot >= 0 && slot < len)
      dst[slot] += amount;
    else
      *stray = 1;
  }
}
// flat index of `slot` (possibly out of range) of node `node` in a [n_nodes_total, len] array, Python-style wrap of
// negative indices; -1 when it falls outside the tensor (torch.put_ raises there)
__device__ __forceinline__ int64_t flat_slot(int64_t node, int len, int slot, int64_t total) {
  int64_t f = node * len + slot;
  if (f < 0) f += total;
  return (f >= 0 && f < total) ? f : -1;
}

// second pass of a period: orders with an out-of-range lead time -> the element the reference's flat put hits
__global__ void __launch_bounds__(256) stray_orders_kernel(HdpoProblem pb, HdpoStatics st, HdpoAction act, HdpoState nxt) {
  const int S = pb.S, W = pb.W, E = pb.E, Wc = W > 0 ? W : 1;
  const int64_t n_store = static_cast<int64_t>(pb.B) * S * Wc, n_wh = static_cast<int64_t>(pb.B) * W,
                n_ech = static_cast<int64_t>(pb.B) * E;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_store) {
    const float a = act.stores[i];
    const int slot = static_cast<int>(st.lead_times[i]) - 1;
    if (a != 0.f && (slot < 0 || slot >= pb.L)) {
      const int64_t f = flat_slot(i / Wc, pb.L, slot, static_cast<int64_t>(pb.B) * S * pb.L);
      if (f >= 0) atomicAdd(nxt.store + f, a);
    }
  } else if (i < n_store + n_wh) {
    const int64_t j = i - n_store;
    const float a = act.warehouses[j];
    const int slot = static_cast<int>(st.warehouse_lead_times[j]) - 1;
    if (a != 0.f && (slot < 0 || slot >= pb.Lw)) {
      const int64_t f = flat_slot(j, pb.Lw, slot, static_cast<int64_t>(pb.B) * W * pb.Lw);
      if (f >= 0) atomicAdd(nxt.warehouse + f, a);
    }
  } else if (i < n_store + n_wh + n_ech) {
    const int64_t j = i - n_store - n_wh;
    const float a = act.echelons[j];
    const int slot = static_cast<int>(st.echelon_lead_times[j]) - 1;
    if (a != 0.f && (slot < 0 || slot >= pb.Le)) {
      const int64_t f = flat_slot(j, pb.Le, slot, static_cast<int64_t>(pb.B) * E * pb.Le);
      if (f >= 0) atomicAdd(nxt.echelon + f, a);
    }
  }
}

__global__ void __launch_bounds__(128) step_fwd_kernel(HdpoProblem pb, HdpoStatics st, HdpoState cur, HdpoAction act,
                                                       const float* __restrict__ demand, int64_t dsb, int64_t dss,
                                                       HdpoState nxt, float* __restrict__ reward,
                                                       int* __restrict__ bad_flag) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pb.B) return;
  const int S = pb.S, W = pb.W, E = pb.E, L = pb.L, Wc = W > 0 ? W : 1;
  int bad = 0;
  float r = 0.f;
  // ---- stores (environment.py:179-234)
  for (int s = 0; s < S; ++s) {
    const int i = b * S + s;
    const float* inv = cur.store + static_cast<int64_t>(i) * L;
    float* out = nxt.store + static_cast<int64_t>(i) * L;
    const float on_hand = inv[0];
    const float d = demand[b * dsb + s * dss];
    const float raw = on_hand - d;
    const float h = st.holding_costs[i], p = st.underage_costs[i];
    float cost;
    if (pb.maximize_profit)
      cost = -p * fminf(on_hand, d) + h * relu0(raw);
    else
      cost = p * relu0(-raw) + h * relu0(raw);
    r += cost;
    const float post = pb.lost_demand ? relu0(raw) : raw;
    shift_pipeline(inv, out, L, post);
    for (int w = 0; w < Wc; ++w)
      land_order(out, L, act.stores[static_cast<int64_t>(i) * Wc + w], st.lead_times[static_cast<int64_t>(i) * Wc + w],
                 &bad);
  }
  // ---- warehouses (environment.py:236-270)
  float wh_orders_total = 0.f;
  for (int w = 0; w < W; ++w) {
    const int i = b * W + w;
    float drawn = 0.f;
    for (int s = 0; s < S; ++s) drawn += act.stores[(static_cast<int64_t>(b) * S + s) * Wc + w];
    const float* inv = cur.warehouse + static_cast<int64_t>(i) * pb.Lw;
    float* out = nxt.warehouse + static_cast<int64_t>(i) * pb.Lw;
    const float raw = inv[0] - drawn;
    const float aw = act.warehouses[i];
    float cost = st.warehouse_holding_costs[i] * relu0(raw);
    if (pb.has_edge_cost) cost += st.warehouse_edge_costs[i] * aw;
    r += cost;
    wh_orders_total += aw;
    shift_pipeline(inv, out, pb.Lw, raw);
    land_order(out, pb.Lw, aw, st.warehouse_lead_times[i], &bad);
  }
  // ---- extra echelons (environment.py:272-299): echelon e is drawn by echelon e+1, the last by the warehouses
  for (int e = 0; e < E; ++e) {
    const int i = b * E + e;
    const float drawn = (e + 1 < E) ? act.echelons[i + 1] : wh_orders_total;
    const float* inv = cur.echelon + static_cast<int64_t>(i) * pb.Le;
    float* out = nxt.echelon + static_cast<int64_t>(i) * pb.Le;
    const float raw = inv[0] - drawn;
    r += st.echelon_holding_costs[i] * relu0(raw);
    shift_pipeline(inv, out, pb.Le, raw);
    land_order(out, pb.Le, act.echelons[i], st.echelon_lead_times[i], &bad);
  }
  reward[b] = r;
  if (bad && bad_flag) atomicOr(bad_flag, 1);
}

// Adjoint (SURVEY.md section 8a recurrences; torch sub-gradient conventions: clip passes gradient at 0,
// minimum ties split 1/2, exact-zero allocations get no pipeline gradient).
__global__ void __launch_bounds__(128) step_bwd_kernel(HdpoProblem pb, HdpoStatics st, HdpoState cur, HdpoAction act,
                                                       const float* __restrict__ demand, int64_t dsb, int64_t dss,
                                                       HdpoState g_next, const float* __restrict__ g_reward,
                                                       HdpoState g_cur, HdpoAction g_act) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pb.B) return;
  const int S = pb.S, W = pb.W, E = pb.E, L = pb.L, Wc = W > 0 ? W : 1;
  const float rb = g_reward ? g_reward[b] : 0.f;
  float g_raw_w[kMaxNodes];
  float wh_orders_total = 0.f;
  for (int w = 0; w < W; ++w) wh_orders_total += act.warehouses[b * W + w];
  // ---- echelons first: their draw-down adjoint feeds the warehouse / echelon order adjoints
  float g_raw_last_echelon = 0.f;
  float g_raw_prev = 0.f;
  for (int e = 0; e < E; ++e) {
    const int i = b * E + e;
    const int Le = pb.Le;
    const float drawn = (e + 1 < E) ? act.echelons[i + 1] : wh_orders_total;
    const float raw = cur.echelon[static_cast<int64_t>(i) * Le] - drawn;
    const float* gn = g_next.echelon ? g_next.echelon + static_cast<int64_t>(i) * Le : nullptr;
    const float gn0 = gn ? gn[0] : 0.f;
    const float g_raw = rb * st.echelon_holding_costs[i] * ge0(raw) + gn0;
    float* gc = g_cur.echelon + static_cast<int64_t>(i) * Le;
    gc[0] = g_raw;
    gc[1] = gn0;
    for (int k = 2; k < Le; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
    const float a = act.echelons[i];
    float ga = 0.f;
    if (a != 0.f && gn) {
      const int slot = static_cast<int>(st.echelon_lead_times[i]) - 1;
      if (slot >= 0 && slot < Le) {
        ga = gn[slot];
      } else {  // stray order: the adjoint of the element the flat put hit
        const int64_t f = flat_slot(i, Le, slot, static_cast<int64_t>(pb.B) * E * Le);
        if (f >= 0) ga = g_next.echelon[f];
      }
    }
    if (e >= 1) ga -= g_raw_prev;  // this echelon's order drew down echelon e-1
    g_act.echelons[i] = ga;
    g_raw_prev = g_raw;
    if (e == E - 1) g_raw_last_echelon = g_raw;
  }
  // ---- warehouses
  for (int w = 0; w < W; ++w) {
    const int i = b * W + w;
    const int Lw = pb.Lw;
    float drawn = 0.f;
    for (int s = 0; s < S; ++s) drawn += act.stores[(static_cast<int64_t>(b) * S + s) * Wc + w];
    const float raw = cur.warehouse[static_cast<int64_t>(i) * Lw] - drawn;
    const float* gn = g_next.warehouse ? g_next.warehouse + static_cast<int64_t>(i) * Lw : nullptr;
    const float gn0 = gn ? gn[0] : 0.f;
    const float g_raw = rb * st.warehouse_holding_costs[i] * ge0(raw) + gn0;
    if (w < kMaxNodes) g_raw_w[w] = g_raw;
    float* gc = g_cur.warehouse + static_cast<int64_t>(i) * Lw;
    gc[0] = g_raw;
    gc[1] = gn0;
    for (int k = 2; k < Lw; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
    const float a = act.warehouses[i];
    float ga = 0.f;
    if (a != 0.f && gn) {
      const int slot = static_cast<int>(st.warehouse_lead_times[i]) - 1;
      if (slot >= 0 && slot < Lw) {
        ga = gn[slot];
      } else {
        const int64_t f = flat_slot(i, Lw, slot, static_cast<int64_t>(pb.B) * W * Lw);
        if (f >= 0) ga = g_next.warehouse[f];
      }
    }
    if (pb.has_edge_cost) ga += rb * st.warehouse_edge_costs[i];
    if (E > 0) ga -= g_raw_last_echelon;
    g_act.warehouses[i] = ga;
  }
  // ---- stores
  for (int s = 0; s < S; ++s) {
    const int i = b * S + s;
    const float on_hand = cur.store[static_cast<int64_t>(i) * L];
    const float d = demand[b * dsb + s * dss];
    const float raw = on_hand - d;
    const float h = st.holding_costs[i], p = st.underage_costs[i];
    const float* gn = g_next.store ? g_next.store + static_cast<int64_t>(i) * L : nullptr;
    const float gn0 = gn ? gn[0] : 0.f;
    float g0;
    if (pb.maximize_profit) {
      const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
      g0 = rb * (-p * tie + h * ge0(raw));
    } else {
      g0 = rb * (-p * le0(raw) + h * ge0(raw));
    }
    g0 += pb.lost_demand ? gn0 * ge0(raw) : gn0;
    float* gc = g_cur.store + static_cast<int64_t>(i) * L;
    gc[0] = g0;
    gc[1] = gn0;
    for (int k = 2; k < L; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
    for (int w = 0; w < Wc; ++w) {
      const int64_t j = static_cast<int64_t>(i) * Wc + w;
      const float a = act.stores[j];
      float ga = 0.f;
      if (a != 0.f && gn) {
        const int slot = static_cast<int>(st.lead_times[j]) - 1;
        if (slot >= 0 && slot < L) {
          ga = gn[slot];
        } else {
          const int64_t f = flat_slot(i, L, slot, static_cast<int64_t>(pb.B) * S * L);
          if (f >= 0) ga = g_next.store[f];
        }
      }
      if (W > 0) ga -= g_raw_w[w];
      g_act.stores[j] = ga;
    }
  }
}

// ---- warp-per-scenario forms -------------------------------------------------------------------------------
constexpr int kStepWarps = 4;           // scenarios (warps) per CTA
constexpr int kStepWarpMinStores = 4;   // fewer stores: the thread-per-scenario kernels below

__device__ __forceinline__ float step_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kStepWarps * 32) step_fwd_warp_kernel(HdpoProblem pb, HdpoStatics st, HdpoState cur,
                                                                        HdpoAction act, const float* __restrict__ demand,
                                                                        int64_t dsb, int64_t dss, HdpoState nxt,
                                                                        float* __restrict__ reward,
                                                                        int* __restrict__ bad_flag) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kStepWarps + (threadIdx.x >> 5);
  if (b >= pb.B) return;  // (whole warps leave together)
  const int S = pb.S, W = pb.W, E = pb.E, L = pb.L, Wc = W > 0 ? W : 1;
  int bad = 0;
  float r = 0.f;
  // ---- stores (environment.py:179-234): lane = store
  for (int s = lane; s < S; s += 32) {
    const int64_t i = static_cast<int64_t>(b) * S + s;
    const float* inv = cur.store + i * L;
    float* out = nxt.store + i * L;
    const float on_hand = inv[0];
    const float d = demand[b * dsb + s * dss];
    const float raw = on_hand - d;
    const float h = st.holding_costs[i], p = st.underage_costs[i];
    r += pb.maximize_profit ? (-p * fminf(on_hand, d) + h * relu0(raw)) : (p * relu0(-raw) + h * relu0(raw));
    shift_pipeline(inv, out, L, pb.lost_demand ? relu0(raw) : raw);
    for (int w = 0; w < Wc; ++w) land_order(out, L, act.stores[i * Wc + w], st.lead_times[i * Wc + w], &bad);
  }
  // ---- warehouses (environment.py:236-270): the draw-down of warehouse w is a warp sum, its update runs on lane w
  float my_orders = 0.f;
  for (int w = 0; w < W; ++w) {
    float part = 0.f;
    for (int s = lane; s < S; s += 32) part += act.stores[(static_cast<int64_t>(b) * S + s) * Wc + w];
    const float drawn = step_warp_sum(part);
    if (lane == (w & 31)) {
      const int64_t i = static_cast<int64_t>(b) * W + w;
      const float* inv = cur.warehouse + i * pb.Lw;
      float* out = nxt.warehouse + i * pb.Lw;
      const float raw = inv[0] - drawn;
      const float aw = act.warehouses[i];
      float cost = st.warehouse_holding_costs[i] * relu0(raw);
      if (pb.has_edge_cost) cost += st.warehouse_edge_costs[i] * aw;
      r += cost;
      my_orders += aw;
      shift_pipeline(inv, out, pb.Lw, raw);
      land_order(out, pb.Lw, aw, st.warehouse_lead_times[i], &bad);
    }
  }
  // ---- extra echelons (environment.py:272-299) on lane 0: echelon e is drawn by echelon e+1, the last by the warehouses
  if (E > 0) {
    const float wh_orders_total = step_warp_sum(my_orders);
    if (lane == 0) {
      for (int e = 0; e < E; ++e) {
        const int64_t i = static_cast<int64_t>(b) * E + e;
        const float drawn = (e + 1 < E) ? act.echelons[i + 1] : wh_orders_total;
        const float* inv = cur.echelon + i * pb.Le;
        float* out = nxt.echelon + i * pb.Le;
        const float raw = inv[0] - drawn;
        r += st.echelon_holding_costs[i] * relu0(raw);
        shift_pipeline(inv, out, pb.Le, raw);
        land_order(out, pb.Le, act.echelons[i], st.echelon_lead_times[i], &bad);
      }
    }
  }
  r = step_warp_sum(r);
  if (lane == 0) reward[b] = r;
  if (bad && bad_flag) atomicOr(bad_flag, 1);
}

__global__ void __launch_bounds__(kStepWarps * 32) step_bwd_warp_kernel(HdpoProblem pb, HdpoStatics st, HdpoState cur,
                                                                        HdpoAction act, const float* __restrict__ demand,
                                                                        int64_t dsb, int64_t dss, HdpoState g_next,
                                                                        const float* __restrict__ g_reward,
                                                                        HdpoState g_cur, HdpoAction g_act) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kStepWarps + (threadIdx.x >> 5);
  if (b >= pb.B) return;
  const int S = pb.S, W = pb.W, E = pb.E, L = pb.L, Wc = W > 0 ? W : 1;
  const float rb = g_reward ? g_reward[b] : 0.f;
  // ---- echelons on lane 0 (their draw-down adjoint feeds the warehouse / echelon order adjoints)
  float g_raw_last_echelon = 0.f;
  if (E > 0) {
    if (lane == 0) {
      float wh_orders_total = 0.f;
      for (int w = 0; w < W; ++w) wh_orders_total += act.warehouses[static_cast<int64_t>(b) * W + w];
      float g_raw_prev = 0.f;
      for (int e = 0; e < E; ++e) {
        const int64_t i = static_cast<int64_t>(b) * E + e;
        const int Le = pb.Le;
        const float drawn = (e + 1 < E) ? act.echelons[i + 1] : wh_orders_total;
        const float raw = cur.echelon[i * Le] - drawn;
        const float* gn = g_next.echelon ? g_next.echelon + i * Le : nullptr;
        const float gn0 = gn ? gn[0] : 0.f;
        const float g_raw = rb * st.echelon_holding_costs[i] * ge0(raw) + gn0;
        float* gc = g_cur.echelon + i * Le;
        gc[0] = g_raw;
        gc[1] = gn0;
        for (int k = 2; k < Le; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
        const float a = act.echelons[i];
        float ga = 0.f;
        if (a != 0.f && gn) {
          const int slot = static_cast<int>(st.echelon_lead_times[i]) - 1;
          if (slot >= 0 && slot < Le) {
            ga = gn[slot];
          } else {  // stray order: the adjoint of the element the flat put hit
            const int64_t f = flat_slot(i, Le, slot, static_cast<int64_t>(pb.B) * E * Le);
            if (f >= 0) ga = g_next.echelon[f];
          }
        }
        if (e >= 1) ga -= g_raw_prev;  // this echelon's order drew down echelon e-1
        g_act.echelons[i] = ga;
        g_raw_prev = g_raw;
        if (e == E - 1) g_raw_last_echelon = g_raw;
      }
    }
    g_raw_last_echelon = __shfl_sync(0xffffffffu, g_raw_last_echelon, 0);
  }
  // ---- warehouses: every lane needs g_raw of every warehouse (its stores' orders drew it down); lane w writes
  float g_raw_w[kMaxNodes];
#pragma unroll
  for (int w = 0; w < kMaxNodes; ++w) {
    g_raw_w[w] = 0.f;
    if (w < W) {
      const int64_t i = static_cast<int64_t>(b) * W + w;
      const int Lw = pb.Lw;
      float part = 0.f;
      for (int s = lane; s < S; s += 32) part += act.stores[(static_cast<int64_t>(b) * S + s) * Wc + w];
      const float drawn = step_warp_sum(part);
      const float raw = cur.warehouse[i * Lw] - drawn;
      const float* gn = g_next.warehouse ? g_next.warehouse + i * Lw : nullptr;
      const float gn0 = gn ? gn[0] : 0.f;
      const float g_raw = rb * st.warehouse_holding_costs[i] * ge0(raw) + gn0;
      g_raw_w[w] = g_raw;
      if (lane == (w & 31)) {
        float* gc = g_cur.warehouse + i * Lw;
        gc[0] = g_raw;
        gc[1] = gn0;
        for (int k = 2; k < Lw; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
        const float a = act.warehouses[i];
        float ga = 0.f;
        if (a != 0.f && gn) {
          const int slot = static_cast<int>(st.warehouse_lead_times[i]) - 1;
          if (slot >= 0 && slot < Lw) {
            ga = gn[slot];
          } else {
            const int64_t f = flat_slot(i, Lw, slot, static_cast<int64_t>(pb.B) * W * Lw);
            if (f >= 0) ga = g_next.warehouse[f];
          }
        }
        if (pb.has_edge_cost) ga += rb * st.warehouse_edge_costs[i];
        if (E > 0) ga -= g_raw_last_echelon;
        g_act.warehouses[i] = ga;
      }
    }
  }
  // ---- stores: lane = store
  for (int s = lane; s < S; s += 32) {
    const int64_t i = static_cast<int64_t>(b) * S + s;
    const float on_hand = cur.store[i * L];
    const float d = demand[b * dsb + s * dss];
    const float raw = on_hand - d;
    const float h = st.holding_costs[i], p = st.underage_costs[i];
    const float* gn = g_next.store ? g_next.store + i * L : nullptr;
    const float gn0 = gn ? gn[0] : 0.f;
    float g0;
    if (pb.maximize_profit) {
      const float tie = on_hand < d ? 1.f : (on_hand == d ? 0.5f : 0.f);
      g0 = rb * (-p * tie + h * ge0(raw));
    } else {
      g0 = rb * (-p * le0(raw) + h * ge0(raw));
    }
    g0 += pb.lost_demand ? gn0 * ge0(raw) : gn0;
    float* gc = g_cur.store + i * L;
    gc[0] = g0;
    gc[1] = gn0;
    for (int k = 2; k < L; ++k) gc[k] = gn ? gn[k - 1] : 0.f;
#pragma unroll
    for (int w = 0; w < kMaxNodes; ++w) {
      if (w < Wc) {
        const int64_t j = i * Wc + w;
        const float a = act.stores[j];
        float ga = 0.f;
        if (a != 0.f && gn) {
          const int slot = static_cast<int>(st.lead_times[j]) - 1;
          if (slot >= 0 && slot < L) {
            ga = gn[slot];
          } else {
            const int64_t f = flat_slot(i, L, slot, static_cast<int64_t>(pb.B) * S * L);
            if (f >= 0) ga = g_next.store[f];
          }
        }
        if (W > 0) ga -= g_raw_w[w];
        g_act.stores[j] = ga;
      }
    }
  }
}

__global__ void allocation_shift_kernel(int64_t* __restrict__ shift, int B, int n, int len) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * n) return;
  const int64_t b = i / n, s = i % n;
  shift[i] = b * (static_cast<int64_t>(len) * n) + s * len;
}

// dst[i][:] = src[idx[i]][:]  (batch assembly on the device: replaces DataLoader's per-sample collate + H2D copy)
__global__ void __launch_bounds__(256) gather_rows_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                          const int64_t* __restrict__ idx, int64_t n_rows,
                                                          int64_t row_floats, int vec4) {
  const int64_t per_row = vec4 ? row_floats / 4 : row_floats;
  const int64_t total = n_rows * per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / per_row, c = i % per_row;
    const int64_t s = idx[r];
    if (vec4)
      reinterpret_cast<float4*>(dst)[r * per_row + c] = reinterpret_cast<const float4*>(src)[s * per_row + c];
    else
      dst[r * per_row + c] = src[s * per_row + c];
  }
}

int validate_problem(const HdpoProblem* pb) {
  HDPO_REQUIRE(pb != nullptr, "null HdpoProblem");
  HDPO_REQUIRE(pb->B >= 0 && pb->S >= 1, "bad B=%d / S=%d", pb->B, pb->S);
  HDPO_REQUIRE(pb->W >= 0 && pb->W <= kMaxNodes, "n_warehouses=%d out of range [0,%d]", pb->W, kMaxNodes);
  HDPO_REQUIRE(pb->E >= 0 && pb->E <= kMaxNodes, "n_extra_echelons=%d out of range [0,%d]", pb->E, kMaxNodes);
  HDPO_REQUIRE(pb->L >= 2, "store pipeline length L=%d must be >= 2 (environment.py:405-412 needs slot 1)", pb->L);
  HDPO_REQUIRE(pb->W == 0 || pb->Lw >= 2, "warehouse pipeline length Lw=%d must be >= 2", pb->Lw);
  HDPO_REQUIRE(pb->E == 0 || pb->Le >= 2, "echelon pipeline length Le=%d must be >= 2", pb->Le);
  HDPO_REQUIRE(pb->E == 0 || pb->W > 0, "extra echelons need at least one warehouse (environment.py:283)");
  return HDPO_OK;
}

// HDPO_STEP_THREAD=1 forces the thread-per-scenario kernels (A/B timing on the GPU box: tools/step_time.py)
static bool step_warp_form(const HdpoProblem* pb) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("HDPO_STEP_THREAD");
    forced = (e && atoi(e) != 0) ? 1 : 0;
  }
  return !forced && pb->S >= kStepWarpMinStores;
}

static int check_step_args(const HdpoProblem* pb, const HdpoStatics* st, const HdpoState* cur, const HdpoAction* act,
                           const float* demand) {
  int rc = validate_problem(pb);
  if (rc) return rc;
  HDPO_REQUIRE(st && cur && act && demand, "null argument");
  HDPO_REQUIRE(st->holding_costs && st->underage_costs && st->lead_times, "store statics missing");
  HDPO_REQUIRE(cur->store && act->stores, "store state / action missing");
  if (pb->W > 0) {
    HDPO_REQUIRE(cur->warehouse && act->warehouses && st->warehouse_lead_times && st->warehouse_holding_costs,
                 "warehouse tensors missing");
    HDPO_REQUIRE(!pb->has_edge_cost || st->warehouse_edge_costs, "has_edge_cost set but warehouse_edge_costs is NULL");
  }
  if (pb->E > 0)
    HDPO_REQUIRE(cur->echelon && act->echelons && st->echelon_lead_times && st->echelon_holding_costs,
                 "echelon tensors missing");
  return HDPO_OK;
}

}  // namespace hdpo

using namespace hdpo;

extern "C" int hdpo_step_fwd(const HdpoProblem* pb, const HdpoStatics* st, const HdpoState* cur, const HdpoAction* act,
                             const float* demand, int64_t dsb, int64_t dss, HdpoState* next, float* reward,
                             void* stream) {
  int rc = check_step_args(pb, st, cur, act, demand);
  if (rc) return rc;
  HDPO_REQUIRE(next && next->store && reward, "null output");
  HDPO_REQUIRE(next->store != cur->store, "next must not alias cur");
  HDPO_REQUIRE(pb->W == 0 || next->warehouse, "next->warehouse missing");
  HDPO_REQUIRE(pb->E == 0 || next->echelon, "next->echelon missing");
  if (pb->B == 0) return HDPO_OK;
  if (step_warp_form(pb)) {
    auto k = step_fwd_warp_kernel;
    HDPO_LAUNCH(k, ceil_div(pb->B, kStepWarps), kStepWarps * 32, 0, stream, *pb, *st, *cur, *act, demand, dsb, dss, *next,
                reward, static_cast<int*>(nullptr));
  } else {
    auto k = step_fwd_kernel;
    HDPO_LAUNCH(k, ceil_div(pb->B, 128), 128, 0, stream, *pb, *st, *cur, *act, demand, dsb, dss, *next, reward,
                static_cast<int*>(nullptr));
  }
  // orders whose lead time is outside [1, pipeline length]: the reference's flat put lands them in a neighbouring node
  const int64_t n_orders = static_cast<int64_t>(pb->B) * (static_cast<int64_t>(pb->S) * (pb->W > 0 ? pb->W : 1) + pb->W + pb->E);
  auto k2 = stray_orders_kernel;
  HDPO_LAUNCH(k2, static_cast<unsigned>(ceil_div64(n_orders, 256)), 256, 0, stream, *pb, *st, *act, *next);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

extern "C" int hdpo_step_bwd(const HdpoProblem* pb, const HdpoStatics* st, const HdpoState* cur, const HdpoAction* act,
                             const float* demand, int64_t dsb, int64_t dss, const HdpoState* g_next,
                             const float* g_reward, HdpoState* g_cur, HdpoAction* g_act, void* stream) {
  int rc = check_step_args(pb, st, cur, act, demand);
  if (rc) return rc;
  HDPO_REQUIRE(g_cur && g_act && g_cur->store && g_act->stores, "null output");
  HDPO_REQUIRE(pb->W == 0 || (g_cur->warehouse && g_act->warehouses), "warehouse adjoint outputs missing");
  HDPO_REQUIRE(pb->E == 0 || (g_cur->echelon && g_act->echelons), "echelon adjoint outputs missing");
  if (pb->B == 0) return HDPO_OK;
  HdpoState gn = {nullptr, nullptr, nullptr};
  if (g_next) gn = *g_next;
  if (step_warp_form(pb)) {
    auto k = step_bwd_warp_kernel;
    HDPO_LAUNCH(k, ceil_div(pb->B, kStepWarps), kStepWarps * 32, 0, stream, *pb, *st, *cur, *act, demand, dsb, dss, gn,
                g_reward, *g_cur, *g_act);
  } else {
    auto k = step_bwd_kernel;
    HDPO_LAUNCH(k, ceil_div(pb->B, 128), 128, 0, stream, *pb, *st, *cur, *act, demand, dsb, dss, gn, g_reward, *g_cur,
                *g_act);
  }
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

extern "C" int hdpo_allocation_shift(int64_t* shift, int32_t B, int32_t n_nodes, int32_t len, void* stream) {
  HDPO_REQUIRE(shift && B >= 0 && n_nodes >= 1 && len >= 1, "bad arguments");
  if (B == 0) return HDPO_OK;
  const int64_t n = static_cast<int64_t>(B) * n_nodes;
  auto k = allocation_shift_kernel;
  HDPO_LAUNCH(k, static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, stream, shift, B, n_nodes, len);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}

extern "C" int hdpo_gather_rows(float* dst, const float* src, const int64_t* idx, int64_t n_rows, int64_t row_floats,
                                void* stream) {
  HDPO_REQUIRE(dst && src && idx && n_rows >= 0 && row_floats >= 1, "bad arguments");
  if (n_rows == 0) return HDPO_OK;
  const int vec4 = (row_floats % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % 16 == 0);
  const int64_t total = n_rows * (vec4 ? row_floats / 4 : row_floats);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto k = gather_rows_kernel;
  HDPO_LAUNCH(k, static_cast<unsigned>(blocks), 256, 0, stream, dst, src, idx, n_rows, row_floats, vec4);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}
