"""Synthetic instances of the BASELINE.json configurations, generated on the device.

Each builder returns (pspec, problem_params, data dict of CUDA tensors in the reference layouts, layer shapes).
Values follow the shipped YAMLs (one_store_backlogged.yml, serial_system.yml, ...); only the demand traces are
synthetic draws of the same distributions (no dataset files are involved).
"""
import math

import torch

from .engine import PolicySpec

WORKLOADS = {}


def register(name):
    def deco(fn):
        WORKLOADS[name] = fn
        return fn
    return deco


def init_flat_params(widths, gen, device):
    """torch.nn.Linear default init (U(+-1/sqrt(fan_in)) for weight and bias), flattened in state_dict order."""
    chunks = []
    for i in range(len(widths) - 1):
        k = 1.0 / math.sqrt(widths[i])
        chunks.append((torch.rand(widths[i + 1] * widths[i], generator=gen, device=device) * 2 - 1) * k)
        chunks.append((torch.rand(widths[i + 1], generator=gen, device=device) * 2 - 1) * k)
    return torch.cat(chunks).contiguous()


def _one_store(B, T, L, lead, dist, lost, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    if dist == "poisson":
        dem = torch.poisson(torch.full((B, 1, T), 5.0, device=device), generator=g)
    else:
        dem = (torch.randn(B, 1, T, generator=g, device=device) * 1.6 + 5.0).clamp_(min=0)
    data = {
        "demands": dem,
        "initial_inventories": 5.0 * torch.rand(B, 1, L, generator=g, device=device),
        "holding_costs": torch.ones(B, 1, device=device),
        "underage_costs": torch.full((B, 1), 9.0, device=device),
        "lead_times": torch.full((B, 1, 1), float(lead), device=device),
    }
    pp = {"n_stores": 1, "n_warehouses": 0, "n_extra_echelons": 0, "lost_demand": lost, "maximize_profit": False}
    widths = [L, 32, 32, 32, 1]
    return PolicySpec("vanilla_one_store", (widths, "elu", None)), pp, data, widths


@register("one_store_lost")
def one_store_lost(device, B=8192, T=50, seed=57):
    """cfg 1: one_store_lost.yml + vanilla_one_store.yml (Poisson(5), lead 4, p=9, h=1)."""
    return _one_store(B, T, 4, 4, "poisson", True, device, seed)


@register("one_store_backlogged_lead20")
def one_store_backlogged_lead20(device, B=1 << 20, T=50, seed=57):
    """cfg 2: one_store_backlogged.yml with the longest tested lead time (20), N(5,1.6) clipped at 0."""
    return _one_store(B, T, 20, 20, "normal", False, device, seed)


@register("serial_system")
def serial_system(device, B=1 << 20, T=50, seed=57):
    """cfg 3: serial_system.yml + vanilla_serial.yml (4 stages: 2 extra echelons, warehouse, store)."""
    g = torch.Generator(device=device).manual_seed(seed)
    dem = (torch.randn(B, 1, T, generator=g, device=device) * 2.0 + 5.0).clamp_(min=0)
    data = {
        "demands": dem,
        "initial_inventories": 5.0 * torch.rand(B, 1, 4, generator=g, device=device),
        "holding_costs": torch.ones(B, 1, device=device),
        "underage_costs": torch.full((B, 1), 9.0, device=device),
        "lead_times": torch.full((B, 1, 1), 4.0, device=device),
        "initial_warehouse_inventories": torch.zeros(B, 1, 3, device=device),
        "warehouse_lead_times": torch.full((B, 1), 3.0, device=device),
        "warehouse_holding_costs": torch.full((B, 1), 0.5, device=device),
        "initial_echelon_inventories": torch.zeros(B, 2, 4, device=device),
        "echelon_lead_times": torch.tensor([2.0, 4.0], device=device).expand(B, 2).contiguous(),
        "echelon_holding_costs": torch.tensor([0.1, 0.2], device=device).expand(B, 2).contiguous(),
    }
    pp = {"n_stores": 1, "n_warehouses": 1, "n_extra_echelons": 2, "lost_demand": False, "maximize_profit": False}
    widths = [15, 32, 32, 4]
    return PolicySpec("vanilla_serial", (widths, "elu", None), warehouse_upper_bound=20.0), pp, data, widths


def net_list(widths):
    """[(name, widths)] of a workload's nets in state_dict order (a plain list = the single 'master' net)."""
    if isinstance(widths, dict):
        return [(m, widths[m]) for m in ("context", "store", "warehouse")]
    return [("master", widths)]


def init_params(widths, gen, device):
    """Flat parameter vector of every net of the workload, state_dict order."""
    return torch.cat([init_flat_params(w, gen, device) for _, w in net_list(widths)]).contiguous()


def forward_macs(widths, n_stores=1):
    """MLP multiply-accumulates per scenario-period of the forward pass. SymmetryAware uses the FACTORED count of
    SURVEY.md 8d: the context half of the first store / warehouse layer is computed once per scenario."""
    dense = lambda w: sum(w[i] * w[i + 1] for i in range(len(w) - 1))  # noqa: E731
    if not isinstance(widths, dict):
        return dense(widths)
    ctx, st, wh = widths["context"], widths["store"], widths["warehouse"]
    C = ctx[-1]
    local = lambda w: (w[0] - C) * w[1] + dense(w[1:])  # noqa: E731
    return dense(ctx) + C * (st[1] + wh[1]) + n_stores * local(st) + local(wh)


def flops_per_scenario_period(widths, first_layer_dgrad=True, n_stores=1):
    """Algorithmic FLOPs: 2*3*MACs_fwd (forward + dgrad + wgrad); recompute is not counted."""
    macs = forward_macs(widths, n_stores)
    f = 2 * 3 * macs
    if not first_layer_dgrad:
        f -= 2 * widths[0] * widths[1]
    return f


def _many_stores(device, B, T, S, W, seed, adjacency=None, lead=None, wh_holding=(0.3,), wh_edge=None, hidden=512):
    """Shared builder of the warehouse settings (one_warehouse_lost_demand.yml / many_warehouses_lost_demand.yml):
    per-store means U[2.5,7.5], cv U[.25,.5], one-factor correlation 0.5, holding U[.7,1.3], underage U[6.3,11.7]."""
    g = torch.Generator(device=device).manual_seed(seed)
    u = lambda lo, hi, *shape: lo + (hi - lo) * torch.rand(*shape, generator=g, device=device)  # noqa: E731
    mean = u(2.5, 7.5, S)
    std = mean * u(0.25, 0.5, S)
    rho = 0.5
    z0 = torch.randn(B, 1, T, generator=g, device=device)
    zs = torch.randn(B, S, T, generator=g, device=device)
    dem = (mean[None, :, None] + std[None, :, None] * (rho ** 0.5 * z0 + (1 - rho) ** 0.5 * zs)).clamp_(min=0)
    if lead is None:
        lead = torch.randint(2, 4, (S, 1), generator=g, device=device).float().expand(S, W)
    L = int(max(3, lead.max().item()))
    data = {
        "demands": dem.contiguous(),
        "initial_inventories": (mean[None, :, None] * torch.rand(B, S, L, generator=g, device=device)).contiguous(),
        "holding_costs": u(0.7, 1.3, S).expand(B, S).contiguous(),
        "underage_costs": u(6.3, 11.7, S).expand(B, S).contiguous(),
        "lead_times": lead.float().expand(B, S, W).contiguous(),
        "mean": mean.expand(B, S).contiguous(),
        "std": std.expand(B, S).contiguous(),
        "initial_warehouse_inventories": torch.zeros(B, W, 3, device=device),
        "warehouse_lead_times": torch.full((B, W), 3.0, device=device),
        "warehouse_holding_costs": torch.tensor(list(wh_holding), device=device).expand(B, W).contiguous(),
    }
    if wh_edge is not None:
        data["warehouse_edge_costs"] = torch.tensor(list(wh_edge), device=device).expand(B, W).contiguous()
    pp = {"n_stores": S, "n_warehouses": W, "n_extra_echelons": 0, "lost_demand": True, "maximize_profit": False}
    if adjacency is not None:
        pp["warehouse_store_adjacency"] = adjacency
    out = S * W + W
    widths = [S * L + W * 3, hidden, hidden, hidden, out]
    wub = 4.0 * float(mean.sum())
    pspec = PolicySpec("vanilla_warehouse", (widths, "elu", None), warehouse_upper_bound=wub, adjacency=adjacency)
    return pspec, pp, data, widths


@register("one_warehouse_lost_demand")
def one_warehouse_lost_demand(device, B=8192, T=50, seed=57):
    """cfg 4: one_warehouse_lost_demand.yml at 50 stores + vanilla_warehouse.yml (153 -> 512^3 -> 51)."""
    return _many_stores(device, B, T, 50, 1, seed)


@register("many_warehouses_lost_demand")
def many_warehouses_lost_demand(device, B=1024, T=50, seed=57):
    """cfg 5: 3 warehouses x 50 stores, Bernoulli(0.7) adjacency with every store connected, leads in [1,7)
    (SURVEY.md 8d), 1024 scenarios per GPU (8192 over 8 GPUs)."""
    S, W = 50, 3
    g = torch.Generator().manual_seed(7)
    adj = (torch.rand(W, S, generator=g) < 0.7).int()
    for s in range(S):
        if adj[:, s].sum() == 0:
            adj[int(torch.randint(W, (1,), generator=g)), s] = 1
    lead = (torch.randint(1, 7, (S, W), generator=g) * adj.t()).float().to(device)
    return _many_stores(device, B, T, S, W, seed, adjacency=adj.tolist(), lead=lead, wh_holding=(0.3, 0.4, 0.2),
                        wh_edge=(0.5, 1.5, 0.7))


@register("one_warehouse_lost_demand_symmetry_aware")
def one_warehouse_lost_demand_symmetry_aware(device, B=8192, T=50, seed=57):
    """cfg 4 as BASELINE.json names it: one_warehouse_lost_demand.yml at 50 stores + the weight-duplicated
    SymmetryAware policy with this repo's symmetry_aware.yml widths (context 153->256->256 sigmoid, store
    (3+4+256)->32->32->1 softplus applied to every store, warehouse (3+256)->16->16->1 sigmoid)."""
    pspec, pp, data, w = _many_stores(device, B, T, 50, 1, seed)
    C = 256
    L = data["initial_inventories"].shape[2]
    widths = {"context": [w[0], 256, C], "store": [L + 4 + C, 32, 32, 1], "warehouse": [3 + C, 16, 16, 1]}
    sym = PolicySpec("symmetry_aware", (widths["context"], "elu", "sigmoid"),
                     warehouse_upper_bound=pspec.warehouse_upper_bound,
                     store_net=(widths["store"], "elu", "softplus"), warehouse_net=(widths["warehouse"], "elu", "sigmoid"),
                     prop_eps=1e-15)
    return sym, pp, data, widths
