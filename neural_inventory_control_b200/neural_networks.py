"""Policy networks with the reference's interface (neural_networks.py) + the hook the fused kernels need.

Every policy is an ordinary `nn.Module` (same `state_dict` keys `net.<module>.<2i>.{weight,bias}`, same
LazyLinear materialisation order => same initial weights for the same torch seed) whose `forward(observation)`
is the GENERIC path: plain torch ops feeding `Simulator.step` (one K3 kernel per period). Policies that the
fused rollout kernels implement additionally expose `fusable_spec()`; `Trainer.simulate_batch` then runs all T
periods (policy MLP + feasibility projection + simulator + cost) in one forward and one adjoint launch.
"""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from .engine import PolicySpec


class MyNeuralNetwork(nn.Module):
    """Base class: builds `self.net` (ModuleDict of sequential MLPs) from nn_params (neural_networks.py:6-106)."""

    def __init__(self, args, device="cpu"):
        super().__init__()
        self.device = device
        self.trainable = True  # benchmark policies without parameters set this to False
        self.gradient_clipping_norm_value = args.get("gradient_clipping_norm_value", None)
        self.activation_functions = {
            "relu": nn.ReLU(), "elu": nn.ELU(), "tanh": nn.Tanh(), "softmax": nn.Softmax(dim=1),
            "softplus": nn.Softplus(), "sigmoid": nn.Sigmoid(),
        }
        self.warehouse_upper_bound = 0
        self.layers = {}
        self.nn_args = args
        self.net = self.create_module_dict(args)
        if args["initial_bias"] is not None:
            for key, val in args["initial_bias"].items():
                if val is not None:
                    # the last Linear sits before the output activation when there is one
                    self.initialize_bias(key, -2 if args["output_layer_activation"][key] else -1, val)

    def forward(self, observation):
        raise NotImplementedError

    def create_module_dict(self, args):
        return nn.ModuleDict({
            key: self.create_sequential_net(key, args["inner_layer_activations"][key],
                                            args["output_layer_activation"][key],
                                            args["neurons_per_hidden_layer"][key], args["output_sizes"][key])
            for key in args["output_sizes"]})

    def create_sequential_net(self, name, inner_layer_activations, output_layer_activation, neurons_per_hidden_layer,
                              output_size):
        """[LazyLinear, act] per hidden layer, then a Linear (Lazy when there is no hidden layer) [+ output act]."""
        layers = []
        for width in neurons_per_hidden_layer:
            layers += [nn.LazyLinear(width), self.activation_functions[inner_layer_activations]]
        if neurons_per_hidden_layer:
            layers.append(nn.Linear(neurons_per_hidden_layer[-1], output_size))
        else:
            layers.append(nn.LazyLinear(output_size))
        if output_layer_activation is not None:
            layers.append(self.activation_functions[output_layer_activation])
        self.layers[name] = layers
        return nn.Sequential(*layers)

    def initialize_bias(self, key, pos, value):
        self.layers[key][pos].bias.data.fill_(value)

    # ---- feasibility projections (neural_networks.py:111-166)
    def apply_proportional_allocation(self, desired_allocations, available_inventory, transshipment=False):
        if available_inventory.dim() > 1:
            available_inventory = available_inventory.sum(dim=1)
        scale = available_inventory / (desired_allocations.sum(dim=1) + 1e-10)
        if not transshipment:
            scale = torch.clip(scale, max=1.0)
        return desired_allocations * scale[:, None]

    def apply_softmax_feasibility_function(self, store_intermediate_outputs, warehouse_inventory, transshipment=False):
        on_hand = warehouse_inventory[:, :, 0].sum(dim=1)
        logits = store_intermediate_outputs
        if not transshipment:  # constant "hold at the warehouse" logit of 1.0, dropped after the softmax
            logits = torch.cat((logits, torch.ones_like(logits[:, :1])), dim=1)
        share = torch.softmax(logits, dim=1)
        if not transshipment:
            share = share[:, :-1]
        return share * on_hand[:, None]

    def flatten_then_concatenate_tensors(self, tensor_list, dim=1):
        return torch.cat([t.flatten(start_dim=dim) for t in tensor_list], dim=dim)

    def concatenate_signal_to_object_state_tensor(self, object_state, signal):
        return torch.cat((object_state, signal.unsqueeze(1).expand(-1, object_state.size(1), -1)), dim=2)

    def unpack_args(self, args, keys):
        return [args[key] for key in keys] if len(keys) > 1 else args[keys[0]]

    # ---- fused-kernel hook
    def fusable_spec(self):
        """PolicySpec when the fused rollout kernels implement this policy, else None (generic per-step path)."""
        return None

    def _mlp_tuple(self, module):
        """(widths, hidden_act, out_act) of a materialised sequential net, or None if it is still lazy."""
        linears = [m for m in self.net[module] if isinstance(m, nn.Linear)]
        if any(isinstance(m, nn.LazyLinear) for m in self.net[module]):
            return None
        widths = [linears[0].in_features] + [m.out_features for m in linears]
        return widths, self.nn_args["inner_layer_activations"][module], self.nn_args["output_layer_activation"][module]

    def _wub_value(self):
        w = self.warehouse_upper_bound
        return float(w.reshape(-1)[0]) if torch.is_tensor(w) else float(w)


class VanillaOneStore(MyNeuralNetwork):
    """One store, no warehouse: order = softplus(MLP(inventory pipeline) + 1) (neural_networks.py:195-214)."""

    def forward(self, observation):
        x = observation["store_inventories"].flatten(start_dim=1)
        return {"stores": F.softplus(self.net["master"](x) + 1).unsqueeze(2)}

    def fusable_spec(self):
        m = self._mlp_tuple("master")
        if m is None or m[0][-1] != 1:
            return None
        return PolicySpec("vanilla_one_store", m)


class VanillaSerial(MyNeuralNetwork):
    """Serial system: sigmoid(MLP(all pipelines)) x upstream on-hand (top echelon x upper bound)
    (neural_networks.py:314-355). The MLP input is detached, exactly as the reference's torch.tensor(x) does."""

    def forward(self, observation):
        store, wh, ech = (observation[k] for k in ("store_inventories", "warehouse_inventories", "echelon_inventories"))
        E = ech.size(1)
        x = self.flatten_then_concatenate_tensors([store, wh, ech]).detach()
        y = self.net["master"](x)
        bound = torch.cat((self.warehouse_upper_bound.unsqueeze(1).expand(ech.shape[0], -1), ech[:, :, 0], wh[:, :, 0]),
                          dim=1)
        alloc = torch.sigmoid(y) * bound
        return {"stores": alloc[:, -1:].unsqueeze(2), "warehouses": alloc[:, -2:-1].unsqueeze(2),
                "echelons": alloc[:, :E].unsqueeze(2)}

    def fusable_spec(self):
        m = self._mlp_tuple("master")
        if m is None:
            return None
        return PolicySpec("vanilla_serial", m, warehouse_upper_bound=self._wub_value())


class VanillaWarehouse(MyNeuralNetwork):
    """One or many warehouses: per-warehouse masked softmax (+ hold logit) x warehouse on-hand for the stores,
    sigmoid x upper bound for the warehouse orders (neural_networks.py:358-427)."""

    def __init__(self, args, scenario=None, device="cpu"):
        super().__init__(args, device)
        self.scenario = scenario
        self.transshipment = args.get("transshipment", False)

    def _adjacency(self, n_warehouses, n_stores, device):
        if n_warehouses == 1:
            return torch.ones(1, n_stores, device=device)
        adj = self.scenario.problem_params.get("warehouse_store_adjacency", None)
        if adj is None:
            raise ValueError(f"warehouse_store_adjacency matrix required for n_warehouses={n_warehouses}")
        return torch.tensor(adj, dtype=torch.float32, device=device)

    def forward(self, observation):
        store, wh = observation["store_inventories"], observation["warehouse_inventories"]
        S, W = store.size(1), wh.size(1)
        y = self.net["master"](torch.cat((store.flatten(start_dim=1), wh.flatten(start_dim=1)), dim=1))
        adj = self._adjacency(W, S, store.device)
        logits = y[:, :S * W].view(-1, S, W)
        stores = torch.zeros_like(logits)
        for w in range(W):
            connected = adj[w].nonzero(as_tuple=True)[0]
            if len(connected) > 0:
                stores[:, connected, w] = self.apply_softmax_feasibility_function(
                    logits[:, connected, w], wh[:, w:w + 1], transshipment=self.transshipment)
        warehouses = torch.sigmoid(y[:, S * W:]) * self.warehouse_upper_bound
        return {"stores": stores, "warehouses": warehouses.unsqueeze(2)}

    def fusable_spec(self):
        m = self._mlp_tuple("master")
        if m is None:
            return None
        adj = None
        if self.scenario is not None:
            adj = self.scenario.problem_params.get("warehouse_store_adjacency", None)
        return PolicySpec("vanilla_warehouse", m, warehouse_upper_bound=self._wub_value(), adjacency=adj,
                          transshipment=self.transshipment)


class SymmetryAware(MyNeuralNetwork):
    """Weight-duplicated policy: a context net over the whole state, ONE store net applied to every store
    (local pipeline + mean/std/underage/lead time + context) and a warehouse net (pipeline + context); store
    outputs go through proportional allocation. The class is absent from the reference snapshot's sources and was
    recovered from its stale bytecode; epsilon 1e-15 is the recovered value."""

    prop_eps = 1e-15

    def forward(self, observation):
        store, wh = observation["store_inventories"], observation["warehouse_inventories"]
        feats = torch.stack([observation["mean"], observation["std"], observation["underage_costs"],
                             observation["lead_times"][:, :, 0]], dim=2)
        context = self.net["context"](self.flatten_then_concatenate_tensors([store, wh]))
        wh_out = self.net["warehouse"](self.concatenate_signal_to_object_state_tensor(wh, context))[:, :, 0]
        st_in = self.concatenate_signal_to_object_state_tensor(torch.cat([store, feats], dim=2), context)
        st_out = self.net["store"](st_in)[:, :, 0]
        scale = torch.clip(wh[:, :, 0].sum(dim=1) / (st_out.sum(dim=1) + self.prop_eps), max=1)
        return {"stores": (st_out * scale[:, None]).unsqueeze(2),
                "warehouses": (wh_out * self.warehouse_upper_bound.unsqueeze(1)).unsqueeze(2)}

    def fusable_spec(self):
        nets = [self._mlp_tuple(k) for k in ("context", "store", "warehouse")]
        if any(n is None for n in nets):
            return None
        return PolicySpec("symmetry_aware", nets[0], warehouse_upper_bound=self._wub_value(), store_net=nets[1],
                          warehouse_net=nets[2], prop_eps=self.prop_eps)


class GNN(MyNeuralNetwork):
    """The reference's shipped weight-shared policy (neural_networks.py:742-1492): node / edge MLPs shared by all nodes
    and edges of the supply graph, `num_message_passing` rounds (1 for warehouse-stores networks, E + 1 for the serial
    system), softplus edge outputs, proportional allocation per supplying node. Same modules (`initial_node`,
    `initial_edge`, `node_update`, `edge_update`, `output` => same state_dict keys) and the same arithmetic; the graph is
    turned ONCE into index tensors (edge order: internal edges in row-major adjacency order, supplier edges, demand
    edges, self-loops) and every per-edge Python loop of the reference becomes one gather / index_add. Runs on the
    generic per-step path (torch policy + one K3 kernel per period)."""

    def __init__(self, args, scenario, device="cpu"):
        super().__init__(args, device)
        self.scenario = scenario
        self.transshipment = args.get("transshipment", False)
        self._graph = {}

    # ---- graph structure (neural_networks.py:757-845, 1015-1077), built once per device
    def graph(self, device):
        key = str(device)
        if key in self._graph:
            return self._graph[key]
        pp = self.scenario.problem_params
        S, W, E = pp.get("n_stores", 0), pp.get("n_warehouses", 0), pp.get("n_extra_echelons", 0)
        n = E + W + S
        adj = torch.zeros(n, n)
        supplier, demand = torch.zeros(n), torch.zeros(n)
        if E > 0:  # serial network: echelon_0 -> ... -> warehouse -> store
            supplier[0] = 1
            for i in range(E - 1):
                adj[i, i + 1] = 1
            adj[E - 1, E] = 1
            adj[E, E + W] = 1
            demand[E + W] = 1
        else:  # warehouses -> stores
            supplier[:W] = 1
            demand[W:] = 1
            if W == 1:
                adj[0, W:] = 1
            else:
                wsa = pp.get("warehouse_store_adjacency", None)
                if wsa is None:
                    raise ValueError(f"Multiple warehouses ({W}) detected but no 'warehouse_store_adjacency' matrix found "
                                     "in problem_params. Please specify which stores connect to which warehouses.")
                adj[:W, W:] = torch.tensor(wsa, dtype=torch.float32)
        internal = adj.nonzero(as_tuple=False)  # row-major: the reference's edge order
        src, tgt = internal[:, 0], internal[:, 1]
        sup_nodes = supplier.nonzero(as_tuple=False).squeeze(-1)
        dem_nodes = demand.nonzero(as_tuple=False).squeeze(-1)
        self_nodes = (adj.sum(dim=1) > 0).nonzero(as_tuple=False).squeeze(-1)
        if self.transshipment:
            self_nodes = self_nodes[:0]
        n_int, n_sup, n_dem, n_self = len(src), len(sup_nodes), len(dem_nodes), len(self_nodes)
        # per edge: source node (n = the virtual all-zero node), target node
        e_src = torch.cat([src, torch.full((n_sup,), n), dem_nodes, self_nodes]).long()
        e_tgt = torch.cat([tgt, sup_nodes, torch.full((n_dem,), n), self_nodes]).long()
        # aggregation: which edges add into incoming[node] / outgoing[node] (neural_networks.py:1196-1268)
        in_edges = torch.cat([torch.arange(n_int), n_int + torch.arange(n_sup), n_int + n_sup + n_dem + torch.arange(n_self)])
        in_nodes = torch.cat([tgt, sup_nodes, self_nodes])
        out_edges = torch.cat([torch.arange(n_int), n_int + n_sup + torch.arange(n_dem),
                               n_int + n_sup + n_dem + torch.arange(n_self)])
        out_nodes = torch.cat([src, dem_nodes, self_nodes])
        in_deg = adj.sum(dim=0) + supplier
        out_deg = adj.sum(dim=1) + demand
        in_deg[self_nodes] += 1
        out_deg[self_nodes] += 1
        in_deg = torch.where(in_deg > 0, in_deg, torch.ones_like(in_deg))
        out_deg = torch.where(out_deg > 0, out_deg, torch.ones_like(out_deg))
        # proportional allocation: the constrained edges of every supplying node = its internal edges + its self-loop
        alloc_edges = torch.cat([torch.arange(n_int), n_int + n_sup + n_dem + torch.arange(n_self)])
        alloc_nodes = torch.cat([src, self_nodes])
        # output mapping (neural_networks.py:1079-1138)
        if E > 0:
            maps = {"stores": [[n_int - 1]], "warehouses": [[n_int - 2]],
                    "echelons": [[n_int]] + [[i - 1] for i in range(1, E)]}
        else:
            stores = [[] for _ in range(S)]
            for i, (a, b) in enumerate(internal.tolist()):
                if a < W <= b:
                    stores[b - W].append(i)
            maps = {"stores": stores, "warehouses": [[n_int + w] for w in range(W)]}
        out_maps = {}
        for k, rows in maps.items():
            if not rows:
                continue
            cols = max(len(r) for r in rows)
            idx = torch.full((len(rows), cols), -1, dtype=torch.long)
            for i, r in enumerate(rows):
                idx[i, :len(r)] = torch.tensor(r, dtype=torch.long)
            out_maps[k] = idx
        g = {"n": n, "S": S, "W": W, "E": E, "src": src, "tgt": tgt, "sup_nodes": sup_nodes, "e_src": e_src, "e_tgt": e_tgt,
             "n_edges": n_int + n_sup + n_dem + n_self, "in_edges": in_edges, "in_nodes": in_nodes, "out_edges": out_edges,
             "out_nodes": out_nodes, "in_scale": 1.0 / torch.sqrt(in_deg), "out_scale": 1.0 / torch.sqrt(out_deg),
             "alloc_edges": alloc_edges, "alloc_nodes": alloc_nodes, "steps": E + 1 if E > 0 else 1, "out_maps": out_maps}
        g = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in g.items()}
        g["out_maps"] = {k: v.to(device) for k, v in out_maps.items()}
        self._graph[key] = g
        return g

    # ---- node features, padded to a common [inventory | state] layout (neural_networks.py:847-911, 1170-1177)
    def node_features(self, obs, g):
        blocks, inv_lens = [], []
        if g["E"] > 0:
            blocks.append([obs["echelon_inventories"], obs["echelon_holding_costs"].unsqueeze(-1)])
            inv_lens.append(obs["echelon_inventories"].size(-1))
        wh = [obs["warehouse_inventories"], obs["warehouse_holding_costs"].unsqueeze(-1)]
        if "warehouse_edge_costs" in obs and obs["warehouse_edge_costs"] is not None:
            wh.append(obs["warehouse_edge_costs"].unsqueeze(-1))
        blocks.append(wh)
        inv_lens.append(obs["warehouse_inventories"].size(-1))
        st = [obs["store_inventories"], obs["holding_costs"].unsqueeze(-1), obs["underage_costs"].unsqueeze(-1)]
        if "past_demands" in obs:
            st.append(obs["past_demands"])
            if "days_from_christmas" in obs:
                st.append(obs["days_from_christmas"].unsqueeze(-1))
        else:
            st += [obs["mean"].unsqueeze(-1), obs["std"].unsqueeze(-1)]
        blocks.append(st)
        inv_lens.append(obs["store_inventories"].size(-1))
        max_inv = max(inv_lens)
        max_state = max(sum(t.size(-1) for t in b[1:]) for b in blocks)
        padded = []
        for b, il in zip(blocks, inv_lens):
            state = torch.cat(b[1:], dim=-1)
            padded.append(torch.cat([F.pad(b[0], (0, max_inv - il)), F.pad(state, (0, max_state - state.size(-1)))], dim=2))
        return torch.cat(padded, dim=1)

    def edge_lead_times(self, obs, g):
        """[n_edges] lead time feature of every edge, taken from scenario 0 like the reference (:949-1013)."""
        n_int = len(g["src"])
        lt = torch.zeros(g["n_edges"], device=g["src"].device)
        if g["E"] > 0:
            E = g["E"]
            vals = [obs["echelon_lead_times"][0, i + 1] for i in range(E - 1)]
            vals += [obs["warehouse_lead_times"][0, 0], obs["lead_times"][0, 0, 0]]
            lt[:n_int] = torch.stack(vals)
            lt[n_int] = obs["echelon_lead_times"][0, 0]
        else:
            W = g["W"]
            lt[:n_int] = obs["lead_times"][0][g["tgt"] - W, g["src"]]
            lt[n_int:n_int + len(g["sup_nodes"])] = obs["warehouse_lead_times"][0, g["sup_nodes"]]
        return lt

    def _gather_nodes(self, nodes, idx):
        """nodes[:, idx] with index n = the virtual supplier / customer node (all zeros)."""
        ext = torch.cat([nodes, torch.zeros_like(nodes[:, :1])], dim=1)
        return ext[:, idx]

    def forward(self, observation):
        dev = observation["store_inventories"].device
        g = self.graph(dev)
        feats = self.node_features(observation, g)
        nodes = self.net["initial_node"](feats)
        B = nodes.size(0)
        lt = self.edge_lead_times(observation, g).view(1, -1, 1).expand(B, -1, -1)
        edges = self.net["initial_edge"](torch.cat([self._gather_nodes(nodes, g["e_src"]),
                                                    self._gather_nodes(nodes, g["e_tgt"]), lt], dim=-1))
        for _ in range(g["steps"]):
            incoming = torch.zeros(B, g["n"], edges.size(-1), device=dev, dtype=edges.dtype)
            outgoing = torch.zeros_like(incoming)
            incoming = incoming.index_add(1, g["in_nodes"], edges[:, g["in_edges"]]) * g["in_scale"].view(1, -1, 1)
            outgoing = outgoing.index_add(1, g["out_nodes"], edges[:, g["out_edges"]]) * g["out_scale"].view(1, -1, 1)
            nodes = nodes + self.net["node_update"](torch.cat([nodes, incoming, outgoing], dim=-1))
            edges = edges + self.net["edge_update"](torch.cat([edges, self._gather_nodes(nodes, g["e_src"]),
                                                               self._gather_nodes(nodes, g["e_tgt"])], dim=-1))
        out = self.net["output"](edges).squeeze(-1)  # [B, n_edges]
        # proportional allocation per supplying node over its internal edges + self-loop (:1437-1492, :111-138)
        inv = torch.cat([observation[k][:, :, 0] for k in ("echelon_inventories", "warehouse_inventories",
                                                           "store_inventories") if k in observation and
                         (k != "echelon_inventories" or g["E"] > 0)], dim=1)
        want = torch.zeros(B, g["n"], device=dev, dtype=out.dtype).index_add(1, g["alloc_nodes"], out[:, g["alloc_edges"]])
        scale = inv / (want + 1e-10)
        if not self.transshipment:
            scale = torch.clip(scale, max=1.0)
        alloc = out.clone()
        alloc[:, g["alloc_edges"]] = out[:, g["alloc_edges"]] * scale[:, g["alloc_nodes"]]
        result = {}
        for k, idx in g["out_maps"].items():
            picked = alloc[:, idx.clamp(min=0).reshape(-1)].view(B, *idx.shape)
            result[k] = picked * (idx >= 0).to(picked.dtype)
        return result


class BaseStock(MyNeuralNetwork):
    """order = max(base level - inventory position, 0) (neural_networks.py:216-229); generic path only."""

    def forward(self, observation):
        inv_pos = observation["store_inventories"].sum(dim=2)
        level = self.net["master"](torch.tensor([0.0]).to(self.device))
        return {"stores": torch.clip(level - inv_pos, min=0).unsqueeze(2)}


class CappedBaseStock(MyNeuralNetwork):
    """Base stock with a cap on the order (neural_networks.py:296-311); generic path only."""

    def forward(self, observation):
        inv_pos = observation["store_inventories"].sum(dim=2)
        out = self.net["master"](torch.tensor([0.0]).to(self.device))
        level, cap = out[0], out[1]
        return {"stores": torch.clip(level - inv_pos, min=torch.tensor([0.0]).to(self.device), max=cap).unsqueeze(2)}


class EchelonStock(MyNeuralNetwork):
    """Echelon base-stock policy for the serial system (neural_networks.py:231-294); generic path only."""

    def forward(self, observation):
        store, wh, ech = (observation[k] for k in ("store_inventories", "warehouse_inventories", "echelon_inventories"))
        E = ech.size(1)
        raw = F.softplus(self.net["master"](torch.tensor([0.0]).to(self.device)) + 10.0)
        base_levels = torch.cumsum(raw, dim=0).flip(dims=[0])  # partial sums, upstream first
        inv_pos = torch.cat((ech.sum(dim=2), wh.sum(dim=2), store.sum(dim=2)), dim=1)
        upstream_on_hand = torch.cat((1000000 * torch.ones_like(wh[:, :, 0]), ech[:, :, 0], wh[:, :, 0]), dim=1)
        tentative = torch.clip(torch.stack([base_levels[k] - inv_pos[:, k:].sum(dim=1) for k in range(2 + E)], dim=1),
                               min=0)
        alloc = torch.minimum(tentative, upstream_on_hand)
        return {"stores": alloc[:, -1:].unsqueeze(2), "warehouses": alloc[:, -2:-1].unsqueeze(2),
                "echelons": alloc[:, :E].unsqueeze(2)}


class DataDrivenNet(MyNeuralNetwork):
    """One MLP over inventories + real-data features (past demands, costs, days from Christmas, lead times)
    (neural_networks.py:430-519). With warehouses the net emits W warehouse orders followed by S*W store orders,
    masked by the adjacency and scaled down per warehouse when they exceed what the warehouse holds."""

    def __init__(self, args, scenario=None, device="cpu"):
        super().__init__(args, device)
        self.scenario = scenario
        self._edge_mask = None

    def _mask(self, ref):
        if self._edge_mask is None or self._edge_mask.device != ref.device:
            adj = self.scenario.problem_params["warehouse_store_adjacency"]
            self._edge_mask = torch.tensor(adj, dtype=torch.float32, device=ref.device).t().contiguous()  # [S, W]
        return self._edge_mask.to(ref.dtype)

    def forward(self, observation):
        store_inv = observation["store_inventories"]
        has_wh = "warehouse_inventories" in observation and observation["warehouse_inventories"].size(1) > 0
        feats = [store_inv]
        if has_wh:
            feats.append(observation["warehouse_inventories"])
        feats += [observation[k] for k in ("past_demands", "underage_costs", "holding_costs", "days_from_christmas",
                                           "lead_times")]
        out = self.net["master"](self.flatten_then_concatenate_tensors(feats))
        if not has_wh:
            return {"stores": out.unsqueeze(2)}
        wh_inv = observation["warehouse_inventories"]
        B, S, W = out.size(0), store_inv.size(1), wh_inv.size(1)
        wanted = out[:, W:].reshape(B, S, W) * self._mask(out)
        # per warehouse: scale the connected stores' orders by min(1, inventory / total wanted); "inventory" is the
        # whole pipeline row of the warehouse, as in the reference (apply_proportional_allocation sums a 2-D argument)
        scale = torch.clip(wh_inv.sum(dim=2) / (wanted.sum(dim=1) + 1e-10), max=1.0)
        return {"stores": wanted * scale[:, None, :], "warehouses": out[:, :W].unsqueeze(2)}


class QuantilePolicy(MyNeuralNetwork):
    """Policies that choose a demand QUANTILE per store and turn it into a base-stock level with a frozen quantile
    forecaster (neural_networks.py:521-591). Subclasses define compute_desired_quantiles."""

    def __init__(self, args, device="cpu"):
        super().__init__(args=args, device=device)
        self.fixed_nets = {"quantile_forecaster": self.load_forecaster(args, requires_grad=False)}
        self.allow_back_orders = False

    def load_forecaster(self, nn_params, requires_grad=True):
        from .quantile_forecaster import FullyConnectedForecaster
        import numpy as np
        net = FullyConnectedForecaster([128, 128], lead_times=nn_params["forecaster_lead_times"],
                                       qs=np.arange(0.05, 1, 0.05))
        # the shipped checkpoint was written from a CUDA process: map it to wherever this policy lives
        net.load_state_dict(torch.load(f"{nn_params['forecaster_location']}", map_location="cpu"))
        for p in net.parameters():
            p.requires_grad_(requires_grad)
        return net.to(self.device)

    def _apply(self, fn, *a, **k):  # the forecaster is not a registered sub-module (it must stay out of state_dict)
        out = super()._apply(fn, *a, **k)
        for net in getattr(self, "fixed_nets", {}).values():
            net._apply(fn)
        return out

    def forecast_base_stock_allocation(self, past_demands, days_from_christmas, store_inventories, lead_times,
                                       quantiles, allow_back_orders=False):
        feats = torch.cat((past_demands,
                           days_from_christmas.unsqueeze(1).expand(past_demands.shape[0], past_demands.shape[1], 1)),
                          dim=2)
        level = self.fixed_nets["quantile_forecaster"].get_quantile(feats, quantiles, lead_times)
        order = level - store_inventories.sum(dim=2)
        if not allow_back_orders:
            order = torch.clip(order, min=0)
        return {"stores": order.unsqueeze(2)}

    def compute_desired_quantiles(self, args):
        raise NotImplementedError

    def forward(self, observation):
        quantiles = self.compute_desired_quantiles({k: observation[k] for k in ("underage_costs", "holding_costs")})
        return self.forecast_base_stock_allocation(
            observation["past_demands"], observation["days_from_christmas"], observation["store_inventories"],
            observation["lead_times"][:, :, 0], quantiles, allow_back_orders=self.allow_back_orders)


def _newsvendor_ratio(args):
    return args["underage_costs"] / (args["underage_costs"] + args["holding_costs"])


class TransformedNV(QuantilePolicy):
    """Learned monotone-free map of the newsvendor ratio u/(u+h) to a quantile (neural_networks.py:593-600)."""

    def compute_desired_quantiles(self, args):
        return self.net["master"](_newsvendor_ratio(args))


class QuantileNV(QuantilePolicy):
    """Newsvendor quantile u/(u+h) itself; nothing to train (neural_networks.py:602-614)."""

    def __init__(self, args, device="cpu"):
        super().__init__(args=args, device=device)
        self.trainable = False

    def compute_desired_quantiles(self, args):
        return _newsvendor_ratio(args)


class ReturnsNV(QuantileNV):
    """QuantileNV that may also return stock (negative orders): a non-admissible benchmark (neural_networks.py:616-625)."""

    def __init__(self, args, device="cpu"):
        super().__init__(args=args, device=device)
        self.allow_back_orders = True


class FixedQuantile(QuantilePolicy):
    """One learned quantile for every store and period (neural_networks.py:627-634)."""

    def compute_desired_quantiles(self, args):
        q = self.net["master"](torch.tensor([0.0]).to(self.device))
        return q.unsqueeze(1).expand(args["underage_costs"].shape[0], args["underage_costs"].shape[1])


class JustInTime(MyNeuralNetwork):
    """Clairvoyant benchmark: orders exactly the demand that will materialise when the order arrives
    (neural_networks.py:637-740). With warehouses every store is served by its connected warehouse with the shortest
    (batch-mean) lead time, and that warehouse orders the store's demand one warehouse lead time further out."""

    def __init__(self, args, scenario=None, device="cpu"):
        super().__init__(args=args, device=device)
        self.scenario = scenario
        self.trainable = False

    @staticmethod
    def _future(demands, when):
        return torch.gather(demands, 2, when.unsqueeze(2)).squeeze(2)

    def forward(self, observation):
        demands, shift = self.unpack_args(observation["internal_data"], ["demands", "period_shift"])
        now = int(observation["current_period"].reshape(-1)[0]) + shift
        B, S, horizon = demands.shape
        lead = observation["lead_times"]
        n_wh = observation["warehouse_inventories"].size(1) if "warehouse_inventories" in observation else 0
        if n_wh == 0:
            when = torch.clip(now + lead[:, :, 0].long(), max=horizon - 1)
            return {"stores": torch.clip(self._future(demands, when), min=0).unsqueeze(2)}
        adj = torch.tensor(self.scenario.problem_params["warehouse_store_adjacency"], dtype=torch.float32,
                           device=lead.device).t()  # [S, W]
        served = adj.sum(dim=1) > 0
        mean_lead = torch.where(adj > 0, lead.mean(dim=0), torch.full_like(adj, float("inf")))
        pick = torch.argmin(mean_lead, dim=1)  # [S] supplying warehouse of each store
        lead_sel = torch.gather(lead, 2, pick[None, :, None].expand(B, S, 1)).squeeze(2).long()
        wh_lead_sel = observation["warehouse_lead_times"][:, pick].long()  # [B, S]
        store_need = self._future(demands, torch.clip(now + lead_sel, max=horizon - 1)) * served
        wh_need = self._future(demands, torch.clip(now + wh_lead_sel + lead_sel, max=horizon - 1)) * served
        stores = torch.zeros(B, S, n_wh, device=lead.device, dtype=demands.dtype)
        stores.scatter_(2, pick[None, :, None].expand(B, S, 1), store_need.unsqueeze(2))
        orders = torch.zeros(B, n_wh, device=lead.device, dtype=demands.dtype).index_add_(1, pick, wh_need)
        return {"stores": torch.clip(stores, min=0), "warehouses": torch.clip(orders, min=0).unsqueeze(2)}


class NeuralNetworkCreator:
    """name -> class registry, default output sizes, warehouse upper bound (neural_networks.py:1495-1574)."""

    def set_default_output_size(self, module_name, problem_params):
        S, W = problem_params["n_stores"], problem_params["n_warehouses"]
        master = S * W + W if W > 1 else S + W
        return {"master": master, "store": 1, "warehouse": 1, "context": None}[module_name]

    def get_architecture(self, name):
        return {
            "vanilla_one_store": VanillaOneStore,
            "base_stock": BaseStock,
            "capped_base_stock": CappedBaseStock,
            "echelon_stock": EchelonStock,
            "vanilla_serial": VanillaSerial,
            "vanilla_warehouse": VanillaWarehouse,
            "symmetry_aware": SymmetryAware,
            "data_driven": DataDrivenNet,
            "transformed_nv": TransformedNV,
            "fixed_quantile": FixedQuantile,
            "quantile_nv": QuantileNV,
            "returns_nv": ReturnsNV,
            "just_in_time": JustInTime,
            "gnn": GNN,
        }[name]

    def get_warehouse_upper_bound(self, warehouse_upper_bound_mult, scenario, device="cpu"):
        mean = scenario.store_params["demand"]["mean"]
        if type(mean) == float:  # noqa: E721 - numpy floats must NOT match, exactly like the reference
            mean = [mean]
        return torch.tensor([warehouse_upper_bound_mult * sum(mean)]).float().to(device)

    def create_neural_network(self, scenario, nn_params, device="cpu"):
        params = copy.deepcopy(nn_params)
        for key, val in params["output_sizes"].items():
            if val is None:
                params["output_sizes"][key] = self.set_default_output_size(key, scenario.problem_params)
        cls = self.get_architecture(params["name"])
        if params["name"] in ("vanilla_warehouse", "gnn", "just_in_time", "data_driven"):
            model = cls(params, scenario, device=device)
        else:
            model = cls(params, device=device)
        if "warehouse_upper_bound_mult" in nn_params.keys():
            model.warehouse_upper_bound = self.get_warehouse_upper_bound(nn_params["warehouse_upper_bound_mult"],
                                                                         scenario, device)
        return model.to(device)
