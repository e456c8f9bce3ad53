"""Differentiable simulator with the reference's interface (environment.py:7-531), backed by the K3 kernels.

`Simulator.reset(periods, problem_params, data, observation_params) -> (observation, None)` and
`Simulator.step(action) -> (observation, reward[B], terminated, None, None)` keep the reference's observation
dict, `_internal_data` (incl. the int64 allocation-shift tables) and period bookkeeping, but one period is ONE
CUDA kernel (hdpo_step_fwd) and its autograd ONE kernel (hdpo_step_bwd) instead of ~60 eager ops and 2-4 host
syncs. This is the generic path for arbitrary torch policies; fusable policies never come here during training
(Trainer.simulate_batch runs the whole rollout in the fused kernels).
"""
import ctypes as C

import torch

from . import _capi as K
from . import _lib, spec
from .engine import _f32c, _ptr, current_stream_ptr


class _StepFn(torch.autograd.Function):
    """(new store inv, new wh inv, new echelon inv, reward) = step(state, action); backward = hdpo_step_bwd."""

    @staticmethod
    def forward(ctx, sim, demand_col, store, wh, ech, a_store, a_wh, a_ech):
        lib, pb, st = sim._lib, sim._pb, sim._statics
        dev = store.device
        store_c, wh_c, ech_c = _f32c(store, "store_inventories"), _f32c(wh, "wh"), _f32c(ech, "ech")
        a_s, a_w, a_e = _f32c(a_store, "action.stores"), _f32c(a_wh, "action.warehouses"), _f32c(a_ech, "action.echelons")
        n_store = torch.empty_like(store_c)
        n_wh = torch.empty_like(wh_c) if wh_c is not None else None
        n_ech = torch.empty_like(ech_c) if ech_c is not None else None
        reward = torch.empty(pb.B, dtype=torch.float32, device=dev)
        cur = K.State(_ptr(store_c), _ptr(wh_c), _ptr(ech_c))
        nxt = K.State(_ptr(n_store), _ptr(n_wh), _ptr(n_ech))
        act = K.Action(_ptr(a_s), _ptr(a_w), _ptr(a_e))
        dem, sb, ss = demand_col
        rc = lib.hdpo_step_fwd(C.byref(pb), C.byref(st), C.byref(cur), C.byref(act), dem, sb, ss, C.byref(nxt),
                               reward.data_ptr(), current_stream_ptr(dev))
        K.check(lib, rc, "hdpo_step_fwd")
        ctx.sim, ctx.demand_col = sim, demand_col
        ctx.saved = (store_c, wh_c, ech_c, a_s, a_w, a_e)
        ctx.mark_non_differentiable()
        outs = (n_store, n_wh if n_wh is not None else store_c.new_empty(0),
                n_ech if n_ech is not None else store_c.new_empty(0), reward)
        return outs

    @staticmethod
    def backward(ctx, g_store, g_wh, g_ech, g_reward):
        sim = ctx.sim
        lib, pb, st = sim._lib, sim._pb, sim._statics
        store_c, wh_c, ech_c, a_s, a_w, a_e = ctx.saved
        dev = store_c.device

        def opt(g, ref):
            if ref is None or g is None:
                return None
            return _f32c(g, "grad")

        gn = K.State(_ptr(opt(g_store, store_c)), _ptr(opt(g_wh, wh_c)), _ptr(opt(g_ech, ech_c)))
        keep = (opt(g_store, store_c), opt(g_wh, wh_c), opt(g_ech, ech_c))  # noqa: F841 - keep alive during the launch
        gr = _f32c(g_reward, "g_reward") if g_reward is not None else None
        gc_store = torch.empty_like(store_c)
        gc_wh = torch.empty_like(wh_c) if wh_c is not None else None
        gc_ech = torch.empty_like(ech_c) if ech_c is not None else None
        ga_s = torch.empty_like(a_s)
        ga_w = torch.empty_like(a_w) if a_w is not None else None
        ga_e = torch.empty_like(a_e) if a_e is not None else None
        cur = K.State(_ptr(store_c), _ptr(wh_c), _ptr(ech_c))
        act = K.Action(_ptr(a_s), _ptr(a_w), _ptr(a_e))
        g_cur = K.State(_ptr(gc_store), _ptr(gc_wh), _ptr(gc_ech))
        g_act = K.Action(_ptr(ga_s), _ptr(ga_w), _ptr(ga_e))
        dem, sb, ss = ctx.demand_col
        rc = lib.hdpo_step_bwd(C.byref(pb), C.byref(st), C.byref(cur), C.byref(act), dem, sb, ss, C.byref(gn),
                               _ptr(gr), C.byref(g_cur), C.byref(g_act), current_stream_ptr(dev))
        K.check(lib, rc, "hdpo_step_bwd")
        return None, None, gc_store, gc_wh, gc_ech, ga_s, ga_w, ga_e


class Simulator:
    """Differentiable inventory simulator (gym-style reset/step; the gym spaces of the reference are inert
    containers nobody reads and are not reproduced - SURVEY.md 2.1 row 21)."""

    metadata = {"render_modes": None}

    def __init__(self, device="cpu"):
        self.device = device
        self.problem_params, self.observation_params, self.maximize_profit = None, None, None
        self.batch_size, self.n_stores, self.periods, self.observation, self._internal_data = None, None, None, None, None
        self.action_space = None
        self.observation_space = None
        self._lib = None

    # ------------------------------------------------------------------ reset
    def reset(self, periods, problem_params, data, observation_params):
        self.problem_params = problem_params
        self.observation_params = observation_params
        self.batch_size, self.n_stores, self.periods = len(data["initial_inventories"]), problem_params["n_stores"], periods
        self._data = data
        self._internal_data = {"demands": data["demands"], "period_shift": observation_params["demand"]["period_shift"]}
        for kind in ("time_features", "sample_features"):
            if observation_params[kind] is not None:
                self._internal_data.update({k: data[k] for k in observation_params[kind]})
        self._lib = _lib.load()
        inv = data["initial_inventories"]
        if not inv.is_cuda:
            raise RuntimeError("Simulator needs CUDA tensors: the HDPO engine has no CPU path "
                               "(the CPU reference lives in oracle/ for tests only)")
        self._internal_data["allocation_shift"] = self.initialize_shifts_for_allocation_put(inv.shape)
        W, E = problem_params["n_warehouses"], problem_params["n_extra_echelons"]
        if W > 0:
            self._internal_data["warehouse_allocation_shift"] = self.initialize_shifts_for_allocation_put(
                data["initial_warehouse_inventories"].shape)
        if E > 0:
            self._internal_data["echelon_allocation_shift"] = self.initialize_shifts_for_allocation_put(
                data["initial_echelon_inventories"].shape)
        self._internal_data["zero_allocation_tensor"] = self.initialize_zero_allocation_tensor(inv.shape[:-1])
        self.observation = self.initialize_observation(data, observation_params)
        self.maximize_profit = problem_params["maximize_profit"]

        B, S, L = inv.shape
        Lw = data["initial_warehouse_inventories"].shape[2] if W > 0 else 0
        Le = data["initial_echelon_inventories"].shape[2] if E > 0 else 0
        has_edge = W > 0 and data.get("warehouse_edge_costs") is not None
        self._pb = spec.problem(B, S, W, E, L, Lw, Le, problem_params["lost_demand"],
                                problem_params["maximize_profit"], has_edge)
        keys = ("holding_costs", "underage_costs", "lead_times", "warehouse_lead_times", "warehouse_holding_costs",
                "warehouse_edge_costs", "echelon_lead_times", "echelon_holding_costs", "mean", "std")
        self._static_tensors = {k: _f32c(data.get(k), k) for k in keys}
        self._statics = K.Statics(*[_ptr(self._static_tensors[k]) for k in keys])
        self._demands = _f32c(data["demands"], "demands")
        return self.observation, None

    def initialize_shifts_for_allocation_put(self, shape):
        """int64 table shift[b,s] = b*(L*S) + s*L (environment.py:77-101), computed on the device."""
        B, n, L = shape
        out = torch.empty(B, n, dtype=torch.int64, device=self.device)
        rc = self._lib.hdpo_allocation_shift(out.data_ptr(), B, n, L, current_stream_ptr(out.device))
        K.check(self._lib, rc, "hdpo_allocation_shift")
        return out

    def initialize_zero_allocation_tensor(self, shape):
        return torch.zeros(shape, device=self.device)

    def initialize_observation(self, data, observation_params):
        obs = {"store_inventories": data["initial_inventories"], "current_period": torch.tensor([0])}
        if observation_params["include_warehouse_inventory"]:
            obs["warehouse_lead_times"] = data["warehouse_lead_times"]
            obs["warehouse_holding_costs"] = data["warehouse_holding_costs"]
            obs["warehouse_inventories"] = data["initial_warehouse_inventories"]
            if data.get("warehouse_edge_costs") is not None:
                obs["warehouse_edge_costs"] = data["warehouse_edge_costs"]
        if self.problem_params["n_extra_echelons"] > 0:
            obs["echelon_lead_times"] = data["echelon_lead_times"]
            obs["echelon_holding_costs"] = data["echelon_holding_costs"]
            obs["echelon_inventories"] = data["initial_echelon_inventories"]
        for k, v in observation_params["include_static_features"].items():
            if v:
                obs[k] = data[k]
        if observation_params["demand"]["past_periods"] > 0:
            obs["past_demands"] = self.update_past_demands(data, observation_params, self.batch_size, self.n_stores, 0)
        if observation_params["time_features"]:
            self.update_time_features(data, obs, observation_params, current_period=0)
        if observation_params["sample_features"] is not None:
            for k in observation_params["sample_features"]:
                obs[k] = data[k]
        return obs

    # ------------------------------------------------------------------ step
    def step(self, action):
        t = int(self.observation["current_period"].item())  # CPU tensor: no device sync
        col = t + self._internal_data["period_shift"]
        dem = self._demands
        if col >= dem.shape[2]:
            raise ValueError("Current period is greater than the number of periods in the data")
        self.update_past_data()
        self.update_time_features(self._internal_data, self.observation, self.observation_params, current_period=t + 1)
        demand_col = (dem.data_ptr() + 4 * col, dem.stride(0), dem.stride(1))
        obs = self.observation
        W, E = self.problem_params["n_warehouses"], self.problem_params["n_extra_echelons"]
        wh = obs["warehouse_inventories"] if W > 0 else None
        ech = obs["echelon_inventories"] if E > 0 else None
        n_store, n_wh, n_ech, reward = _StepFn.apply(self, demand_col, obs["store_inventories"], wh, ech,
                                                     action["stores"], action.get("warehouses") if W > 0 else None,
                                                     action.get("echelons") if E > 0 else None)
        obs["store_inventories"] = n_store
        if W > 0:
            obs["warehouse_inventories"] = n_wh
        if E > 0:
            obs["echelon_inventories"] = n_ech
        obs["current_period"] += 1
        terminated = obs["current_period"] >= self.periods
        return obs, reward, terminated, None, None

    def get_current_demands(self, data, current_period):
        return data["demands"][:, :, current_period + self._internal_data["period_shift"]]

    # ------------------------------------------------------------------ real-data observation features (torch glue)
    def update_past_demands(self, data, observation_params, batch_size, stores, current_period):
        past = observation_params["demand"]["past_periods"]
        now = current_period + self._internal_data["period_shift"]
        if now == 0:
            return torch.zeros(batch_size, stores, past, device=self.device)
        window = data["demands"][:, :, max(0, now - past):now]
        missing = past - window.shape[2]
        if missing > 0:
            window = torch.cat([torch.zeros(batch_size, stores, missing, device=self.device), window], dim=2)
        return window

    def update_time_features(self, data, observation, observation_params, current_period):
        if observation_params["time_features"] is not None:
            for k in observation_params["time_features"]:
                if data[k].shape[2] + 2 < current_period:
                    raise ValueError("Current period is greater than the number of periods in the data")
                col = min(current_period + observation_params["demand"]["period_shift"], data[k].shape[2] - 1)
                observation[k] = data[k][:, :, col]

    def update_past_data(self):
        t = int(self.observation["current_period"].item())
        if self._internal_data["demands"].shape[2] + 2 < t:
            raise ValueError("Current period is greater than the number of periods in the data")
        if self.observation_params["demand"]["past_periods"] > 0:
            self.observation["past_demands"] = self.update_past_demands(
                self._internal_data, self.observation_params, self.batch_size, self.n_stores,
                current_period=min(t + 1, self._internal_data["demands"].shape[2]))
