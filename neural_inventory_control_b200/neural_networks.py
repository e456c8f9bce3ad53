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
        if params["name"] in ("vanilla_warehouse",):
            model = cls(params, scenario, device=device)
        else:
            model = cls(params, device=device)
        if "warehouse_upper_bound_mult" in nn_params.keys():
            model.warehouse_upper_bound = self.get_warehouse_upper_bound(nn_params["warehouse_upper_bound_mult"],
                                                                         scenario, device)
        return model.to(device)
