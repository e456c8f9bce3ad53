"""Host side of the fused rollout: torch tensors in, C-ABI calls on the current CUDA stream, autograd glue.

This is the thin layer the reference-facing `Trainer.simulate_batch` sits on. PyTorch is used for device
memory, streams and autograd bookkeeping only; all arithmetic of the path happens in libhdpo_b200.so.
"""
import ctypes as C

import torch

from . import _capi as K
from . import _lib, spec

_DATA_KEYS = ("holding_costs", "underage_costs", "lead_times", "warehouse_lead_times", "warehouse_holding_costs",
              "warehouse_edge_costs", "echelon_lead_times", "echelon_holding_costs", "mean", "std")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _f32c(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"'{name}' must live on the CUDA device (the HDPO engine has no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def current_stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class PolicySpec:
    """What the kernels need to know about a policy network (built by neural_networks.fusable_spec)."""

    def __init__(self, arch, master, warehouse_upper_bound=0.0, adjacency=None, transshipment=False,
                 store_net=None, warehouse_net=None, prop_eps=1e-15, param_names=None):
        self.arch = arch
        self.master = master  # (widths, hidden_act, out_act)
        self.store_net = store_net
        self.warehouse_net = warehouse_net
        self.warehouse_upper_bound = float(warehouse_upper_bound)
        self.adjacency = adjacency  # None or [W][S] nested list / tensor of 0/1
        self.transshipment = bool(transshipment)
        self.prop_eps = prop_eps
        self.param_names = param_names


class FusedRollout:
    """One (policy, problem, batch-shape) instance of the fused forward rollout + reverse-time adjoint."""

    def __init__(self, pspec, problem_params, data, periods, ignore_periods=0, period_shift=0,
                 discrete_allocation=False, demand_layout=K.DEMAND_BST, precision="fp32", save_for_backward=True,
                 philox=None, checkpoint_interval=0):
        """checkpoint_interval K > 1 (small-net path): tape the state of every K-th period only, the adjoint recomputes.
        philox: None, or {dist: 'normal' | 'poisson', mean: [S] tensor, std: [S] tensor, rho, clip, seed, offset,
        periods: time extent} - the demand trace is then generated on the device inside every forward call (the batch
        dict needs no 'demands'); bump 'offset' per batch for fresh draws (set_philox_offset)."""
        self.lib = _lib.load()
        inv = data["initial_inventories"]
        self.device = inv.device
        if self.device.type != "cuda":
            raise RuntimeError("FusedRollout needs CUDA tensors (no CPU fallback)")
        B, S, L = inv.shape
        W = int(problem_params["n_warehouses"])
        E = int(problem_params["n_extra_echelons"])
        Lw = data["initial_warehouse_inventories"].shape[2] if W > 0 else 0
        Le = data["initial_echelon_inventories"].shape[2] if E > 0 else 0
        self.has_edge = W > 0 and data.get("warehouse_edge_costs") is not None
        self.pb = spec.problem(B, S, W, E, L, Lw, Le, problem_params["lost_demand"],
                               problem_params["maximize_profit"], self.has_edge)
        self.philox = None
        if philox is not None:
            t_stride = int(philox.get("periods", periods + period_shift))
            mean = _f32c(philox["mean"].to(self.device), "mean")
            std = _f32c(philox["std"].to(self.device), "std") if philox.get("std") is not None else None
            self.philox = dict(philox, mean_ptr=mean.data_ptr(), std_ptr=_ptr(std), _keep=(mean, std))
        else:
            dem = data["demands"]
            t_stride = dem.shape[2] if demand_layout == K.DEMAND_BST else dem.shape[0]
        self.adj = None
        if pspec.adjacency is not None and W > 1:
            self.adj = torch.as_tensor(pspec.adjacency, dtype=torch.int32, device=self.device).contiguous()
        self.desc = spec.rollout_desc(pspec.arch, self.pb, periods, t_stride, pspec.master,
                                      period_shift=period_shift, ignore_periods=ignore_periods,
                                      demand_layout=demand_layout, discrete_allocation=discrete_allocation,
                                      transshipment=pspec.transshipment, precision=precision,
                                      save_for_backward=save_for_backward,
                                      warehouse_upper_bound=pspec.warehouse_upper_bound, prop_eps=pspec.prop_eps,
                                      store_net=pspec.store_net, warehouse_net=pspec.warehouse_net,
                                      adjacency_ptr=_ptr(self.adj), philox=self.philox,
                                      checkpoint_interval=checkpoint_interval)
        self.n_params = int(self.lib.hdpo_param_count(C.byref(self.desc)))
        ws = int(self.lib.hdpo_rollout_workspace_bytes(C.byref(self.desc)))
        if ws == 0:
            raise K.HdpoError("no fused rollout for this policy/problem: " + self.lib.hdpo_last_error().decode())
        self.ws_bytes = ws
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
        self.B, self.S, self.T = B, S, periods
        self.cost_b = torch.empty(B, dtype=torch.float32, device=self.device)
        self.report_b = torch.empty(B, dtype=torch.float32, device=self.device)
        self.totals = torch.empty(2, dtype=torch.float64, device=self.device)
        self._keep = None

    def signature(self):
        return (self.B, self.S, self.T, self.desc.t_stride, self.desc.ignore_periods, self.desc.discrete_allocation,
                self.desc.save_for_backward)

    def set_philox_offset(self, offset):
        """Counter offset of the next forward call's demand draws (one batch consumes t_stride * S * B / 4 counters)."""
        self.desc.philox_offset = int(offset)

    def bind(self, data):
        """Resolve the per-batch device pointers (reference layouts, fp32, contiguous)."""
        d = {k: _f32c(data.get(k), k) for k in _DATA_KEYS}
        dem = _f32c(data["demands"], "demands") if self.philox is None else None
        init = [_f32c(data["initial_inventories"], "initial_inventories"),
                _f32c(data.get("initial_warehouse_inventories"), "initial_warehouse_inventories") if self.pb.W else None,
                _f32c(data.get("initial_echelon_inventories"), "initial_echelon_inventories") if self.pb.E else None]
        st = K.Statics(*[_ptr(d[k]) for k in _DATA_KEYS])
        state = K.State(*[_ptr(t) for t in init])
        self._keep = (d, dem, init, st, state)  # keep the tensors alive until backward
        return dem, st, state

    def forward(self, flat_params, data, reward_tb=None, final_state=None):
        dem, st, state = self.bind(data)
        if flat_params.numel() != self.n_params:
            raise ValueError(f"parameter vector has {flat_params.numel()} floats, the descriptor needs {self.n_params}")
        self._params = flat_params
        fin = None
        if final_state is not None:
            fin = K.State(*[_ptr(t) for t in final_state])
        rc = self.lib.hdpo_rollout_fwd(C.byref(self.desc), flat_params.data_ptr(), _ptr(dem), C.byref(st),
                                       C.byref(state), self.cost_b.data_ptr(), self.report_b.data_ptr(),
                                       _ptr(reward_tb), self.totals.data_ptr(),
                                       C.byref(fin) if fin is not None else None, self.workspace.data_ptr(),
                                       self.ws_bytes, current_stream_ptr(self.device))
        K.check(self.lib, rc, "hdpo_rollout_fwd")
        return self.totals

    def backward(self, g_total, g_report=0.0, out=None):
        d, dem, init, st, state = self._keep
        grad = out if out is not None else torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        rc = self.lib.hdpo_rollout_bwd(C.byref(self.desc), self._params.data_ptr(), _ptr(dem), C.byref(st),
                                       float(g_total), float(g_report), grad.data_ptr(), self.workspace.data_ptr(),
                                       self.ws_bytes, current_stream_ptr(self.device))
        K.check(self.lib, rc, "hdpo_rollout_bwd")
        return grad


class _RolloutFn(torch.autograd.Function):
    """(total cost, reported cost) = rollout(params); backward = reverse-time adjoint kernel."""

    @staticmethod
    def forward(ctx, engine, data, flat_params):
        totals = engine.forward(flat_params.detach(), data)
        ctx.engine = engine
        both = totals.to(torch.float32)
        return both[0], both[1]

    @staticmethod
    def backward(ctx, g_total, g_report):
        # one host read of the two upstream scalars (the trainer's own .item() syncs have already happened)
        gt = float(g_total) if g_total is not None else 0.0
        gr = float(g_report) if g_report is not None else 0.0
        grad = ctx.engine.backward(gt, gr)
        return None, None, grad


def rollout(engine, data, flat_params):
    return _RolloutFn.apply(engine, data, flat_params)
