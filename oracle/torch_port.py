"""CPU port of the reference's hot path in PyTorch eager ops - TEST / BASELINE INFRASTRUCTURE ONLY.

Why a second restatement next to the numpy oracle: the reference itself IS PyTorch-eager code running on
the host cores (MKL sgemm, ATen elementwise kernels, autograd). For `bench.py --impl reference` and the
`cpu_baseline` object the honest CPU number is therefore a port that executes the same kind of ATen
work with autograd - not the (slower, single-purpose) numpy oracle. It is written functionally (no
classes, no gym plumbing) and pinned to the same reference-generated goldens
(tests/test_oracle_golden.py::test_torch_port_*), so it is also an independent check of the explicit
adjoint in hdpo_oracle.py.

Only tests/ and bench.py's baseline legs may import this module. Citations are into /root/reference.
"""
import torch
import torch.nn.functional as F

_ACT = {"elu": F.elu, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "softplus": F.softplus,
        None: lambda x: x}


def mlp(x, layers, hidden_act, out_act):
    """layers: list of (weight [out,in], bias [out]) tensors (neural_networks.py:80-106)."""
    n = len(layers)
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        x = _ACT[hidden_act if i < n - 1 else out_act](x)
    return x


def _advance(inv, post, alloc, lead):
    """environment.py:391-434: shift the pipeline one slot and land every non-zero allocation at slot lead-1."""
    L = inv.shape[2]
    head = (post + inv[:, :, 1]).unsqueeze(2)
    new = torch.cat([head, inv[:, :, 2:], torch.zeros_like(head)], dim=2)
    nz = alloc != 0                                   # exact zeros are filtered before the put (no gradient)
    slot = (lead.long() - 1).clamp(0, L - 1)
    return new.scatter_add(2, slot, alloc * nz)


def env_step(pb, state, action, data, t):
    """environment.py:110-169. state/action are dicts; returns (new_state, reward [B])."""
    d = data["demands"][:, :, t + pb["period_shift"]]
    inv = state["store"]
    on_hand = inv[:, :, 0]
    raw = on_hand - d
    hold = data["holding_costs"] * raw.clamp(min=0)
    if pb["maximize_profit"]:
        cost = -data["underage_costs"] * torch.minimum(on_hand, d) + hold
    else:
        cost = data["underage_costs"] * (-raw).clamp(min=0) + hold
    post = raw.clamp(min=0) if pb["lost_demand"] else raw
    new = {"store": _advance(inv, post, action["stores"], data["lead_times"])}
    reward = cost.sum(1)
    if pb["n_warehouses"] > 0:
        wh = state["wh"]
        raw_w = wh[:, :, 0] - action["stores"].sum(1)
        cw = data["warehouse_holding_costs"] * raw_w.clamp(min=0)
        if data.get("warehouse_edge_costs") is not None:
            cw = cw + data["warehouse_edge_costs"] * action["warehouses"].sum(2)
        new["wh"] = _advance(wh, raw_w, action["warehouses"], data["warehouse_lead_times"].unsqueeze(2))
        reward = reward + cw.sum(1)
    if pb["n_extra_echelons"] > 0:
        ech = state["ech"]
        drawn = torch.cat([action["echelons"][:, 1:, :].sum(2), action["warehouses"].sum((1, 2)).unsqueeze(1)], 1)
        raw_e = ech[:, :, 0] - drawn
        new["ech"] = _advance(ech, raw_e, action["echelons"], data["echelon_lead_times"].unsqueeze(2))
        reward = reward + (data["echelon_holding_costs"] * raw_e.clamp(min=0)).sum(1)
    return new, reward


def policy(pol, pb, state, data):
    """pol: dict(arch, layers, hidden_act, out_act, wub, adjacency, transshipment)."""
    store = state["store"]
    B = store.shape[0]
    arch = pol["arch"]
    if arch == "vanilla_one_store":  # neural_networks.py:200-214
        y = mlp(store.flatten(1), pol["layers"], pol["hidden_act"], pol["out_act"])
        return {"stores": F.softplus(y + 1).unsqueeze(2)}
    if arch == "vanilla_serial":  # neural_networks.py:319-355 (input detached by torch.tensor(...))
        wh, ech = state["wh"], state["ech"]
        E = ech.shape[1]
        x = torch.cat([store.flatten(1), wh.flatten(1), ech.flatten(1)], 1).detach()
        y = mlp(x, pol["layers"], pol["hidden_act"], pol["out_act"])
        bound = torch.cat([pol["wub"].reshape(1, 1).expand(B, 1), ech[:, :, 0], wh[:, :, 0]], 1)
        alloc = torch.sigmoid(y) * bound
        return {"echelons": alloc[:, :E].unsqueeze(2), "warehouses": alloc[:, E:E + 1].unsqueeze(2),
                "stores": alloc[:, E + 1:].unsqueeze(2)}
    if arch == "vanilla_warehouse":  # neural_networks.py:369-427
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        y = mlp(torch.cat([store.flatten(1), wh.flatten(1)], 1), pol["layers"], pol["hidden_act"], pol["out_act"])
        logits = y[:, :S * W].view(B, S, W)
        adj = pol["adjacency"] if W > 1 else torch.ones(1, S)
        cols = []
        for w in range(W):
            conn = adj[w].nonzero(as_tuple=True)[0]
            z = logits[:, conn, w]
            if not pol["transshipment"]:
                z = torch.cat([z, torch.ones(B, 1, dtype=z.dtype)], 1)
            p = torch.softmax(z, 1)
            if not pol["transshipment"]:
                p = p[:, :-1]
            col = torch.zeros(B, S, dtype=y.dtype).index_copy(1, conn, p * wh[:, w, 0:1])
            cols.append(col)
        stores = torch.stack(cols, 2)
        return {"stores": stores, "warehouses": (torch.sigmoid(y[:, S * W:]) * pol["wub"]).unsqueeze(2)}
    if arch == "symmetry_aware":  # SURVEY.md 2.3 (recovered forward), neural_networks.py:111-138,168-187
        wh = state["wh"]
        S, W = store.shape[1], wh.shape[1]
        nets = pol["nets"]
        feats = torch.stack([data["mean"], data["std"], data["underage_costs"], data["lead_times"][:, :, 0]], 2)
        ctx = mlp(torch.cat([store.flatten(1), wh.flatten(1)], 1), *nets["context"])
        wo = mlp(torch.cat([wh, ctx.unsqueeze(1).expand(B, W, -1)], 2), *nets["warehouse"])[:, :, 0]
        so = mlp(torch.cat([store, feats, ctx.unsqueeze(1).expand(B, S, -1)], 2), *nets["store"])[:, :, 0]
        scale = torch.clip(wh[:, :, 0].sum(1) / (so.sum(1) + pol.get("prop_eps", 1e-15)), max=1)
        return {"stores": (so * scale[:, None]).unsqueeze(2), "warehouses": (wo * pol["wub"].unsqueeze(1)).unsqueeze(2)}
    raise KeyError(arch)


def simulate(pol, pb, data, T, ignore=0):
    """trainer.py:181-216 with PolicyLoss (reward.sum()). Returns (total, report, reward_tb)."""
    state = {"store": data["initial_inventories"]}
    if pb["n_warehouses"] > 0:
        state["wh"] = data["initial_warehouse_inventories"]
    if pb["n_extra_echelons"] > 0:
        state["ech"] = data["initial_echelon_inventories"]
    total = 0
    report = 0
    rewards = []
    for t in range(T):
        action = policy(pol, pb, state, data)
        state, reward = env_step(pb, state, action, data, t)
        s = reward.sum()
        total = total + s
        if t >= ignore:
            report = report + s
        rewards.append(reward.detach())
    return total, report, torch.stack(rewards, 0)


def train_step(pol, pb, data, T, ignore=0):
    """One batch exactly as trainer.py:163-173: simulate, mean loss, backward. Returns (total, report, grads)."""
    params = [t for wb in pol["layers"] for t in wb]
    for prm in params:
        prm.grad = None
    total, report, _ = simulate(pol, pb, data, T, ignore)
    B = data["demands"].shape[0]
    (total / (B * T * pb["n_stores"])).backward()
    return float(total.detach()), float(report.detach()), [prm.grad for prm in params]
