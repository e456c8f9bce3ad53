"""CPU oracle for the HDPO rollout path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's hot path (policy -> Simulator.step x T -> summed cost ->
backward), written as explicit forward recurrences plus an explicit reverse-time adjoint (no
autograd), so that it documents exactly the arithmetic the CUDA kernels must reproduce.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module, and only as the checker / reported CPU baseline. The product
package `neural_inventory_control_b200` never imports it and has no CPU fallback.

PARITY PIN: the reference ships no tests or golden vectors. This oracle is
pinned instead against outputs of the unmodified reference executed in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; checked by tests/test_oracle_golden.py) for
every function below, in fp32 and against the reference's float64 run.

All `file:line` citations are into the reference tree (/root/reference).
Array conventions: B scenarios, S stores, W warehouses (action width Wc = max(W,1)), E extra
echelons, L / Lw / Le pipeline lengths. dtype follows the inputs (float32 or float64).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# activations (torch semantics, neural_networks.py:36-43)
# ----------------------------------------------------------------------------------------------


def _elu(x):
    # nn.ELU(alpha=1): x if x > 0 else expm1(x)  (ATen uses expm1; verified against torch 2.11)
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def _elu_grad(x, y):
    # elu_backward with the *input*: 1 for x > 0, exp(x) for x <= 0  (== y + 1 there)
    return np.where(x > 0, np.ones_like(x), y + 1)


def _softplus(x):
    # nn.Softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x))
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def _softplus_grad(x):
    z = np.exp(np.minimum(x, 20))
    return np.where(x > 20, np.ones_like(x), z / (z + 1))


def _sigmoid(x):
    e = np.exp(-np.abs(x))  # overflow-free form of 1/(1+exp(-x))
    return np.where(x >= 0, 1 / (1 + e), e / (1 + e))


ACT = {
    "elu": (_elu, lambda x, y: _elu_grad(x, y)),
    "relu": (lambda x: np.maximum(x, 0), lambda x, y: (x > 0).astype(x.dtype)),
    "tanh": (np.tanh, lambda x, y: 1 - y * y),
    "sigmoid": (_sigmoid, lambda x, y: y * (1 - y)),
    "softplus": (_softplus, lambda x, y: _softplus_grad(x)),
    None: (lambda x: x, lambda x, y: np.ones_like(x)),
}

# ----------------------------------------------------------------------------------------------
# MLP  (neural_networks.py:80-106: [LazyLinear, act]*k, Linear [, out_act]; weights [out,in])
# ----------------------------------------------------------------------------------------------


@dataclass
class MLP:
    weights: List[np.ndarray]  # each [out, in]
    biases: List[np.ndarray]   # each [out]
    hidden_act: Optional[str] = "elu"
    out_act: Optional[str] = None

    def astype(self, dt):
        return MLP([w.astype(dt) for w in self.weights], [b.astype(dt) for b in self.biases], self.hidden_act,
                   self.out_act)

    def n_params(self):
        return sum(w.size + b.size for w, b in zip(self.weights, self.biases))


def mlp_forward(net: MLP, x):
    """Returns (y, saved) where saved = list of (layer input, pre-activation, post-activation)."""
    saved = []
    n = len(net.weights)
    for i, (w, b) in enumerate(zip(net.weights, net.biases)):
        z = x @ w.T + b
        act = net.hidden_act if i < n - 1 else net.out_act
        y = ACT[act][0](z)
        saved.append((x, z, y))
        x = y
    return x, saved


def mlp_backward(net: MLP, saved, g_out, need_input_grad=True):
    """Returns (g_input or None, [gW...], [gb...]) for upstream gradient g_out."""
    n = len(net.weights)
    gws, gbs = [None] * n, [None] * n
    g = g_out
    for i in reversed(range(n)):
        x, z, y = saved[i]
        act = net.hidden_act if i < n - 1 else net.out_act
        gz = g * ACT[act][1](z, y)
        gws[i] = gz.T @ x
        gbs[i] = gz.sum(0)
        if i > 0 or need_input_grad:
            g = gz @ net.weights[i]
    return (g if need_input_grad else None), gws, gbs


# ----------------------------------------------------------------------------------------------
# problem / policy description
# ----------------------------------------------------------------------------------------------


@dataclass
class Problem:
    """problem_params + observation_params['demand']['period_shift'] (environment.py:41-50)."""
    n_stores: int
    n_warehouses: int = 0
    n_extra_echelons: int = 0
    lost_demand: bool = True
    maximize_profit: bool = False
    period_shift: int = 0


@dataclass
class Policy:
    arch: str                              # vanilla_one_store | vanilla_serial | vanilla_warehouse | symmetry_aware
    nets: Dict[str, MLP] = field(default_factory=dict)
    warehouse_upper_bound: float = 0.0
    adjacency: Optional[np.ndarray] = None  # [W,S] 0/1 (problem_params['warehouse_store_adjacency'])
    transshipment: bool = False
    prop_eps: float = 1e-15                # symmetry_aware proportional-allocation epsilon


def allocation_shift(B: int, n: int, L: int) -> np.ndarray:
    """environment.py:77-101: shift[b,s] = b*(L*n) + s*L, int64, bit-exact."""
    return (np.arange(B, dtype=np.int64)[:, None] * (L * n) + np.arange(n, dtype=np.int64)[None, :] * L)


# ----------------------------------------------------------------------------------------------
# one simulator period, forward  (environment.py:110-299, 391-434)
# ----------------------------------------------------------------------------------------------


def _pipeline_update(inv, post, alloc, lead):
    """environment.py:391-434. inv [B,n,L], post [B,n], alloc [B,n,k], lead [B,n,k] (float-encoded ints).

    new = [post + inv[...,1], inv[...,2:], 0]; then every alloc != 0 is added at slot lead-1 (flat
    `put(accumulate=True)` at shift+lead-1; summation order = ascending k, as the CPU put does).
    """
    B, n, L = inv.shape
    new = np.empty_like(inv)
    new[:, :, 0] = post + inv[:, :, 1]
    new[:, :, 1:L - 1] = inv[:, :, 2:]
    new[:, :, L - 1] = 0
    flat = new.reshape(-1)
    shift = allocation_shift(B, n, L)
    idx = shift[:, :, None] + lead.astype(np.int64) - 1
    for k in range(alloc.shape[2]):
        a = alloc[:, :, k].reshape(-1)
        nz = a != 0
        np.add.at(flat, idx[:, :, k].reshape(-1)[nz], a[nz])
    return new


def store_step(pb: Problem, inv, demand, a_stores, lead, holding, underage):
    """environment.py:179-234. Returns (new_inv [B,S,L], cost_b [B], raw [B,S])."""
    on_hand = inv[:, :, 0]
    raw = on_hand - demand
    if pb.maximize_profit:
        cost = -underage * np.minimum(on_hand, demand) + holding * np.maximum(raw, 0)
    else:
        cost = underage * np.maximum(-raw, 0) + holding * np.maximum(raw, 0)
    post = np.maximum(raw, 0) if pb.lost_demand else raw
    new = _pipeline_update(inv, post, a_stores, lead)
    return new, cost.sum(1), raw


def warehouse_step(wh_inv, a_stores, a_wh, wh_lead, wh_holding, wh_edge=None):
    """environment.py:236-270. Returns (new_wh_inv, cost_b, raw_w [B,W])."""
    raw = wh_inv[:, :, 0] - a_stores.sum(1)
    cost = wh_holding * np.maximum(raw, 0)
    if wh_edge is not None:
        cost = cost + wh_edge * a_wh.sum(2)
    new = _pipeline_update(wh_inv, raw, a_wh, wh_lead[:, :, None])
    return new, cost.sum(1), raw


def echelon_step(ech_inv, a_ech, a_wh, ech_lead, ech_holding):
    """environment.py:272-299. Echelon e is drawn down by echelon e+1's order, the last by all warehouse orders."""
    sub = np.concatenate([a_ech[:, 1:, :].sum(2), a_wh.sum((1, 2))[:, None]], 1)
    raw = ech_inv[:, :, 0] - sub
    cost = ech_holding * np.maximum(raw, 0)
    new = _pipeline_update(ech_inv, raw, a_ech, ech_lead[:, :, None])
    return new, cost.sum(1), raw


# ----------------------------------------------------------------------------------------------
# policies, forward
# ----------------------------------------------------------------------------------------------


def _softmax_with_hold(logits, hold: bool):
    """neural_networks.py:140-166: softmax over [logits..., 1.0] (constant hold logit), last column dropped."""
    if hold:
        z = np.concatenate([logits, np.ones_like(logits[:, :1])], 1)
    else:
        z = logits
    z = z - z.max(1, keepdims=True)
    e = np.exp(z)
    p = e / e.sum(1, keepdims=True)
    return p[:, :-1] if hold else p


def policy_forward(pol: Policy, pb: Problem, state, static):
    """Returns (action dict, saved-for-backward). state = dict(store, wh, ech); static = data dict."""
    store = state["store"]
    B = store.shape[0]
    dt = store.dtype
    if pol.arch == "vanilla_one_store":  # neural_networks.py:200-214
        x = store.reshape(B, -1)
        y, saved = mlp_forward(pol.nets["master"], x)
        pre = y + 1
        a = _softplus(pre)
        return {"stores": a[:, :, None]}, ("one_store", saved, pre)
    if pol.arch == "vanilla_serial":  # neural_networks.py:319-355
        wh, ech = state["wh"], state["ech"]
        E = ech.shape[1]
        x = np.concatenate([store.reshape(B, -1), wh.reshape(B, -1), ech.reshape(B, -1)], 1)
        y, saved = mlp_forward(pol.nets["master"], x)  # input is DETACHED in the reference (torch.tensor(x))
        bound = np.concatenate([np.full((B, 1), pol.warehouse_upper_bound, dt), ech[:, :, 0], wh[:, :, 0]], 1)
        sg = _sigmoid(y)
        alloc = sg * bound
        act = {"echelons": alloc[:, :E, None], "warehouses": alloc[:, -2:-1, None], "stores": alloc[:, -1:, None]}
        return act, ("serial", saved, sg, bound)
    if pol.arch == "vanilla_warehouse":  # neural_networks.py:369-427
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        x = np.concatenate([store.reshape(B, -1), wh.reshape(B, -1)], 1)
        y, saved = mlp_forward(pol.nets["master"], x)
        logits = y[:, :S * W].reshape(B, S, W)
        adj = np.ones((1, S)) if W == 1 else np.asarray(pol.adjacency)
        alloc = np.zeros((B, S, W), dt)
        probs = []
        for w in range(W):
            conn = np.nonzero(adj[w])[0]
            if len(conn) == 0:
                probs.append(None)
                continue
            p = _softmax_with_hold(logits[:, conn, w], hold=not pol.transshipment)
            alloc[:, conn, w] = p * wh[:, w, 0][:, None]
            probs.append((conn, p))
        sg = _sigmoid(y[:, S * W:])
        a_wh = sg * dt.type(pol.warehouse_upper_bound)
        return {"stores": alloc, "warehouses": a_wh[:, :, None]}, ("warehouse", saved, probs, sg)
    if pol.arch == "symmetry_aware":  # recovered from stale bytecode, SURVEY.md section 2.3
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        feats = np.stack([static["mean"], static["std"], static["underage_costs"], static["lead_times"][:, :, 0]], 2)
        x = np.concatenate([store.reshape(B, -1), wh.reshape(B, -1)], 1)
        ctx, s_ctx = mlp_forward(pol.nets["context"], x)
        C = ctx.shape[1]
        wx = np.concatenate([wh, np.broadcast_to(ctx[:, None, :], (B, W, C))], 2).reshape(B * W, -1)
        wo, s_wh = mlp_forward(pol.nets["warehouse"], wx)
        sx = np.concatenate([store, feats.astype(dt), np.broadcast_to(ctx[:, None, :], (B, S, C))], 2).reshape(B * S, -1)
        so, s_st = mlp_forward(pol.nets["store"], sx)
        so = so.reshape(B, S)
        avail = wh[:, :, 0].sum(1)
        tot = so.sum(1)
        ratio = avail / (tot + dt.type(pol.prop_eps))
        scale = np.minimum(ratio, 1)
        stores = so * scale[:, None]
        a_wh = wo.reshape(B, W) * dt.type(pol.warehouse_upper_bound)
        return ({"stores": stores[:, :, None], "warehouses": a_wh[:, :, None]},
                ("sym", s_ctx, s_wh, s_st, so, avail, tot, ratio, scale, C))
    raise KeyError(pol.arch)


# ----------------------------------------------------------------------------------------------
# rollout forward  (trainer.py:181-216)
# ----------------------------------------------------------------------------------------------


def _initial_state(pb: Problem, data):
    st = {"store": data["initial_inventories"].copy()}
    if pb.n_warehouses > 0:
        st["wh"] = data["initial_warehouse_inventories"].copy()
    if pb.n_extra_echelons > 0:
        st["ech"] = data["initial_echelon_inventories"].copy()
    return st


def env_step(pb: Problem, state, action, data, t):
    """environment.py:110-169 for period t. Returns (new_state, reward_b, saved)."""
    d = data["demands"][:, :, t + pb.period_shift]
    new = {}
    new["store"], r, raw_s = store_step(pb, state["store"], d, action["stores"], data["lead_times"],
                                        data["holding_costs"], data["underage_costs"])
    raw_w = raw_e = None
    if pb.n_warehouses > 0:
        new["wh"], rw, raw_w = warehouse_step(state["wh"], action["stores"], action["warehouses"],
                                              data["warehouse_lead_times"], data["warehouse_holding_costs"],
                                              data.get("warehouse_edge_costs"))
        r = r + rw
    if pb.n_extra_echelons > 0:
        new["ech"], re_, raw_e = echelon_step(state["ech"], action["echelons"], action["warehouses"],
                                              data["echelon_lead_times"], data["echelon_holding_costs"])
        r = r + re_
    return new, r, (d, raw_s, raw_w, raw_e)


def rollout_forward(pol: Policy, pb: Problem, data, T: int, ignore_periods: int = 0, discrete: bool = False,
                    keep_tape: bool = False):
    """Returns dict(reward_tb [T,B], total, report, final state, tape)."""
    state = _initial_state(pb, data)
    rewards, tape = [], []
    for t in range(T):
        action, psaved = policy_forward(pol, pb, state, data)
        if discrete:  # trainer.py:201-202 (torch.round = half-to-even = np.rint)
            action = {k: np.rint(v) for k, v in action.items()}
        new, r, esaved = env_step(pb, state, action, data, t)
        rewards.append(r)
        if keep_tape:
            tape.append((state, action, psaved, esaved))
        state = new
    reward_tb = np.stack(rewards, 0)
    # trainer.py:206-210: per-period reward.sum() accumulated in the working precision
    per_t = reward_tb.sum(1)
    total = per_t.sum()
    report = per_t[ignore_periods:].sum()
    return {"reward_tb": reward_tb, "total": total, "report": report, "final": state, "tape": tape}


# ----------------------------------------------------------------------------------------------
# reverse-time adjoint (what autograd computes for trainer.py:169-173)
# ----------------------------------------------------------------------------------------------


def _pipeline_adjoint(g_new, g_raw, alloc, lead):
    """Adjoint of _pipeline_update: returns (g_inv [B,n,L], g_alloc [B,n,k]) given g_new and g wrt `post`.

    put(accumulate) gradient = gather at the same indices; exact-zero allocations were filtered out
    before the put (environment.py:426-432) and get no gradient.
    """
    B, n, L = g_new.shape
    g_inv = np.zeros_like(g_new)
    g_inv[:, :, 0] = g_raw
    g_inv[:, :, 1] = g_new[:, :, 0]
    g_inv[:, :, 2:] = g_new[:, :, 1:L - 1]
    # gather at the SAME flat indices the put used: shift + lead - 1 over the flattened [B, n, L] tensor, so an order with
    # a lead time outside [1, L] reads the adjoint of the neighbouring node's slot it landed in (index -1 wraps)
    idx = allocation_shift(B, n, L)[:, :, None] + lead.astype(np.int64) - 1
    g_alloc = g_new.reshape(-1)[idx] * (alloc != 0)
    return g_inv, g_alloc


def env_step_adjoint(pb: Problem, state, action, data, esaved, g_new, r_bar):
    """Adjoint of env_step. g_new: dict of adjoints wrt the NEW state; r_bar: dLoss/dreward (scalar).

    Sub-gradient conventions reproduced (verified vs torch 2.11):
    clip(x,min=0) passes gradient at x == 0; minimum ties split 1/2, 1/2.
    """
    d, raw_s, raw_w, raw_e = esaved
    dt = state["store"].dtype
    ge = lambda m: m.astype(dt)  # noqa: E731
    h, p = data["holding_costs"], data["underage_costs"]
    g_state, g_act = {}, {}
    on_hand = state["store"][:, :, 0]
    if pb.maximize_profit:
        g_on_hand_direct = -p * (ge(on_hand < d) + 0.5 * ge(on_hand == d)) * r_bar
        g_raw = r_bar * h * ge(raw_s >= 0)
    else:
        g_on_hand_direct = 0
        g_raw = r_bar * (-p * ge(raw_s <= 0) + h * ge(raw_s >= 0))
    g_post = g_new["store"][:, :, 0]
    g_raw = g_raw + (g_post * ge(raw_s >= 0) if pb.lost_demand else g_post)
    g_state["store"], g_act["stores"] = _pipeline_adjoint(g_new["store"], g_raw + g_on_hand_direct, action["stores"],
                                                          data["lead_times"])
    if pb.n_warehouses > 0:
        g_raw_w = r_bar * data["warehouse_holding_costs"] * ge(raw_w >= 0) + g_new["wh"][:, :, 0]
        g_state["wh"], g_aw = _pipeline_adjoint(g_new["wh"], g_raw_w, action["warehouses"],
                                                data["warehouse_lead_times"][:, :, None])
        g_act["stores"] = g_act["stores"] - g_raw_w[:, None, :]
        if data.get("warehouse_edge_costs") is not None:
            g_aw = g_aw + (r_bar * data["warehouse_edge_costs"])[:, :, None]
        g_act["warehouses"] = g_aw
    if pb.n_extra_echelons > 0:
        g_raw_e = r_bar * data["echelon_holding_costs"] * ge(raw_e >= 0) + g_new["ech"][:, :, 0]
        g_state["ech"], g_ae = _pipeline_adjoint(g_new["ech"], g_raw_e, action["echelons"],
                                                 data["echelon_lead_times"][:, :, None])
        g_ae = g_ae.copy()
        g_ae[:, 1:, 0] -= g_raw_e[:, :-1]
        g_act["echelons"] = g_ae
        g_act["warehouses"] = g_act["warehouses"] - g_raw_e[:, -1][:, None, None]
    return g_state, g_act


def _zero_net_grads(pol: Policy):
    return {k: ([np.zeros_like(w) for w in n.weights], [np.zeros_like(b) for b in n.biases])
            for k, n in pol.nets.items()}


def _acc(gr, name, gws, gbs):
    for i, (gw, gb) in enumerate(zip(gws, gbs)):
        gr[name][0][i] += gw
        gr[name][1][i] += gb


def policy_adjoint(pol: Policy, pb: Problem, state, psaved, g_act, grads):
    """Adds the policy path to the state adjoint; accumulates parameter gradients into `grads`."""
    store = state["store"]
    B = store.shape[0]
    dt = store.dtype
    g_state = {k: np.zeros_like(v) for k, v in state.items()}
    kind = psaved[0]
    if kind == "one_store":
        _, saved, pre = psaved
        g_pre = g_act["stores"][:, :, 0] * _softplus_grad(pre)
        g_x, gws, gbs = mlp_backward(pol.nets["master"], saved, g_pre)
        _acc(grads, "master", gws, gbs)
        g_state["store"] += g_x.reshape(store.shape)
    elif kind == "serial":
        _, saved, sg, bound = psaved
        E = state["ech"].shape[1]
        g_alloc = np.concatenate([g_act["echelons"][:, :, 0], g_act["warehouses"][:, :, 0], g_act["stores"][:, :, 0]], 1)
        g_bound = g_alloc * sg
        g_state["ech"][:, :, 0] += g_bound[:, 1:1 + E]
        g_state["wh"][:, :, 0] += g_bound[:, 1 + E:]
        g_y = g_alloc * bound * sg * (1 - sg)
        _, gws, gbs = mlp_backward(pol.nets["master"], saved, g_y, need_input_grad=False)  # input detached
        _acc(grads, "master", gws, gbs)
    elif kind == "warehouse":
        _, saved, probs, sg = psaved
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        g_logits = np.zeros((B, S, W), dt)
        for w in range(W):
            if probs[w] is None:
                continue
            conn, p = probs[w]
            ga = g_act["stores"][:, conn, w]
            g_state["wh"][:, w, 0] += (ga * p).sum(1)
            gp = ga * wh[:, w, 0][:, None]
            # softmax backward restricted to the kept columns (the hold column has zero upstream gradient)
            g_logits[:, conn, w] = p * (gp - (gp * p).sum(1, keepdims=True))
        g_y = np.concatenate([g_logits.reshape(B, S * W),
                              g_act["warehouses"][:, :, 0] * dt.type(pol.warehouse_upper_bound) * sg * (1 - sg)], 1)
        g_x, gws, gbs = mlp_backward(pol.nets["master"], saved, g_y)
        _acc(grads, "master", gws, gbs)
        nS = store[0].size
        g_state["store"] += g_x[:, :nS].reshape(store.shape)
        g_state["wh"] += g_x[:, nS:].reshape(wh.shape)
    elif kind == "sym":
        _, s_ctx, s_wh, s_st, so, avail, tot, ratio, scale, C = psaved
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        L, Lw = store.shape[2], wh.shape[2]
        eps = dt.type(pol.prop_eps)
        g_stores = g_act["stores"][:, :, 0]
        g_scale = (g_stores * so).sum(1)
        g_so = g_stores * scale[:, None]
        # clip(ratio, max=1): gradient to ratio where ratio <= 1 (clamp passes gradient at the boundary)
        g_ratio = g_scale * (ratio <= 1).astype(dt)
        g_avail = g_ratio / (tot + eps)
        g_tot = -g_ratio * avail / (tot + eps) ** 2
        g_so = g_so + g_tot[:, None]
        g_state["wh"][:, :, 0] += g_avail[:, None]
        g_sx, gws, gbs = mlp_backward(pol.nets["store"], s_st, g_so.reshape(B * S, 1))
        _acc(grads, "store", gws, gbs)
        g_sx = g_sx.reshape(B, S, -1)
        g_state["store"] += g_sx[:, :, :L]
        g_ctx = g_sx[:, :, L + 4:].sum(1)
        g_wo = (g_act["warehouses"][:, :, 0] * dt.type(pol.warehouse_upper_bound)).reshape(B * W, 1)
        g_wx, gws, gbs = mlp_backward(pol.nets["warehouse"], s_wh, g_wo)
        _acc(grads, "warehouse", gws, gbs)
        g_wx = g_wx.reshape(B, W, -1)
        g_state["wh"] += g_wx[:, :, :Lw]
        g_ctx = g_ctx + g_wx[:, :, Lw:].sum(1)
        g_x, gws, gbs = mlp_backward(pol.nets["context"], s_ctx, g_ctx)
        _acc(grads, "context", gws, gbs)
        nS = store[0].size
        g_state["store"] += g_x[:, :nS].reshape(store.shape)
        g_state["wh"] += g_x[:, nS:].reshape(wh.shape)
    else:
        raise KeyError(kind)
    return g_state


def rollout_grad(pol: Policy, pb: Problem, data, T: int, grad_scale: Optional[float] = None):
    """Forward + reverse-time adjoint. Returns (forward result, grads[name] = ([gW], [gb])).

    grad_scale = dLoss/d(total); the trainer uses 1/(B*T*n_stores) (trainer.py:169). Gradient flows
    from ALL T periods; ignore_periods only affects the reported loss (trainer.py:208-210).
    """
    B = data["demands"].shape[0]
    dt = data["initial_inventories"].dtype
    if grad_scale is None:
        grad_scale = 1.0 / (B * T * pb.n_stores)
    r_bar = dt.type(grad_scale)
    fwd = rollout_forward(pol, pb, data, T, keep_tape=True)
    grads = _zero_net_grads(pol)
    g_new = {k: np.zeros_like(v) for k, v in fwd["final"].items()}
    for t in reversed(range(T)):
        state, action, psaved, esaved = fwd["tape"][t]
        g_dyn, g_act = env_step_adjoint(pb, state, action, data, esaved, g_new, r_bar)
        g_pol = policy_adjoint(pol, pb, state, psaved, g_act, grads)
        g_new = {k: g_dyn[k] + g_pol[k] for k in g_dyn}
    fwd["g_initial"] = g_new
    return fwd, grads


# ----------------------------------------------------------------------------------------------
# helpers to build Policy / Problem from the golden fixtures or a torch state_dict
# ----------------------------------------------------------------------------------------------


def mlp_from_state_dict(sd: Dict[str, np.ndarray], module: str, hidden_act, out_act) -> MLP:
    """state_dict keys are net.<module>.<idx>.{weight,bias} with idx = 0,2,4,.. (neural_networks.py:80-106)."""
    idxs = sorted({int(k.split(".")[2]) for k in sd if k.startswith(f"net.{module}.") and k.endswith(".weight")})
    ws = [np.asarray(sd[f"net.{module}.{i}.weight"]) for i in idxs]
    bs = [np.asarray(sd[f"net.{module}.{i}.bias"]) for i in idxs]
    return MLP(ws, bs, hidden_act, out_act)


def flatten_grads(pol: Policy, grads) -> Dict[str, np.ndarray]:
    out = {}
    for name, (gws, gbs) in grads.items():
        for i, (gw, gb) in enumerate(zip(gws, gbs)):
            out[f"net.{name}.{2 * i}.weight"] = gw
            out[f"net.{name}.{2 * i}.bias"] = gb
    return out
