"""TEST INFRASTRUCTURE: call the C ABI (include/hdpo_b200.h) with numpy inputs through one of two backends.

  * CudaBackend - the product library libhdpo_b200.so on cuda:0 (parity tests proper, `-m gpu`)
  * EmuBackend  - the same kernel sources compiled for the host-thread emulator (tests/emu), CPU container

Both go through the ctypes prototypes of neural_inventory_control_b200._capi, i.e. through the C ABI.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

from neural_inventory_control_b200 import _capi as K  # noqa: E402
from neural_inventory_control_b200 import spec  # noqa: E402


class EmuBackend:
    name = "emu"
    _lib = None

    def __init__(self):
        if EmuBackend._lib is None:
            sys.path.insert(0, os.path.join(HERE, "emu"))
            import build_emu
            EmuBackend._lib = K.bind(C.CDLL(build_emu.build()))
        self.lib = EmuBackend._lib
        self.stream = None

    def put(self, a, dtype=np.float32):
        return None if a is None else np.ascontiguousarray(a, dtype=dtype)

    def zeros(self, shape, dtype=np.float32):
        return np.zeros(shape, dtype)

    def ptr(self, h):
        return None if h is None else h.ctypes.data

    def get(self, h):
        return None if h is None else np.array(h)

    def sync(self):
        pass


class CudaBackend:
    name = "cuda"

    def __init__(self):
        import torch
        from neural_inventory_control_b200 import _lib
        self.torch = torch
        self.lib = _lib.load()
        self.dev = torch.device("cuda:0")
        self.stream = None  # legacy default stream; torch's current stream is the default stream in the tests

    def put(self, a, dtype=np.float32):
        if a is None:
            return None
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.dev)

    def zeros(self, shape, dtype=np.float32):
        tdt = {np.float32: self.torch.float32, np.float64: self.torch.float64, np.int64: self.torch.int64,
               np.int32: self.torch.int32, np.uint8: self.torch.uint8}[dtype]
        return self.torch.zeros(shape, dtype=tdt, device=self.dev)

    def ptr(self, h):
        return None if h is None else h.data_ptr()

    def get(self, h):
        return None if h is None else h.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()


class Batch:
    """One batch in the reference layouts, resident on the backend, + the structs pointing at it."""

    def __init__(self, be, pb_meta, data):
        self.be = be
        self.h = {k: be.put(v) for k, v in data.items()}
        h, p = self.h, be.ptr
        B, S, L = data["initial_inventories"].shape
        W, E = pb_meta["n_warehouses"], pb_meta["n_extra_echelons"]
        Lw = data["initial_warehouse_inventories"].shape[2] if W > 0 else 0
        Le = data["initial_echelon_inventories"].shape[2] if E > 0 else 0
        self.edge = data.get("warehouse_edge_costs") is not None
        self.pb = spec.problem(B, S, W, E, L, Lw, Le, pb_meta["lost_demand"], pb_meta["maximize_profit"], self.edge)
        self.st = K.Statics(p(h["holding_costs"]), p(h["underage_costs"]), p(h["lead_times"]),
                            p(h.get("warehouse_lead_times")), p(h.get("warehouse_holding_costs")),
                            p(h.get("warehouse_edge_costs")), p(h.get("echelon_lead_times")),
                            p(h.get("echelon_holding_costs")), p(h.get("mean")), p(h.get("std")))
        self.init = K.State(p(h["initial_inventories"]), p(h.get("initial_warehouse_inventories")),
                            p(h.get("initial_echelon_inventories")))
        self.B, self.S, self.W, self.E, self.L, self.Lw, self.Le = B, S, W, E, L, Lw, Le
        self.Wc = max(W, 1)


def flat_params(params, module="master"):
    idxs = sorted({int(k.split(".")[2]) for k in params if k.startswith(f"net.{module}.") and k.endswith(".weight")})
    chunks, shapes, names = [], [], []
    for i in idxs:
        w, b = params[f"net.{module}.{i}.weight"], params[f"net.{module}.{i}.bias"]
        chunks += [w.ravel(), b.ravel()]
        shapes.append(w.shape)
        names += [(f"net.{module}.{i}.weight", w.shape), (f"net.{module}.{i}.bias", b.shape)]
    return np.concatenate(chunks).astype(np.float32), shapes, names


def unflatten(flat, names):
    out, o = {}, 0
    for name, shape in names:
        n = int(np.prod(shape))
        out[name] = flat[o:o + n].reshape(shape)
        o += n
    return out


def slice_batch(data, n):
    return {k: v[:n] for k, v in data.items()}


def rollout(be, meta, params, data, T=None, ignore=None, demand_layout=K.DEMAND_BST, discrete=False, backward=True,
            g_total=None, g_report=0.0, precision="fp32", philox=None, checkpoint_interval=0):
    """philox: None, or {dist, mean, std, rho, clip, seed, offset, periods}: demands = NULL, generated inside the call."""
    L = be.lib
    pp = meta["problem_params"]
    bt = Batch(be, pp, data)
    T = meta["T"] if T is None else T
    ignore = meta["ignore_periods"] if ignore is None else ignore
    modules = ["context", "store", "warehouse"] if meta["nn_name"] == "symmetry_aware" else ["master"]
    flats, names, nets = [], [], {}
    for m in modules:
        f_m, shapes, names_m = flat_params(params, m)
        flats.append(f_m)
        names += names_m
        nets[m] = (spec.mlp_widths(shapes), meta["inner_layer_activations"][m], meta["output_layer_activation"][m])
    flat = np.concatenate(flats)
    ph = None
    if philox is not None:
        t_stride, dem_h = int(philox["periods"]), None
        mean_h = be.put(np.asarray(philox["mean"], np.float32))
        std_h = be.put(np.asarray(philox["std"], np.float32))
        ph = dict(philox, mean_ptr=be.ptr(mean_h), std_ptr=be.ptr(std_h))
    else:
        dem = np.asarray(data["demands"], np.float32)
        t_stride = dem.shape[2]
        if demand_layout == K.DEMAND_TSB:
            dem = np.ascontiguousarray(dem.transpose(2, 1, 0))
        dem_h = be.put(dem)
    adj = pp.get("warehouse_store_adjacency")
    adj_h = None if adj is None else be.put(np.asarray(adj), np.int32)
    desc = spec.rollout_desc(meta["nn_name"], bt.pb, T, t_stride, nets[modules[0]],
                             store_net=nets.get("store"), warehouse_net=nets.get("warehouse"),
                             prop_eps=meta.get("prop_eps", 1e-15),
                             period_shift=meta.get("period_shift", 0), ignore_periods=ignore,
                             demand_layout=demand_layout, discrete_allocation=discrete,
                             transshipment=meta.get("transshipment", False), save_for_backward=backward,
                             warehouse_upper_bound=meta["warehouse_upper_bound"], adjacency_ptr=be.ptr(adj_h),
                             precision=precision, philox=ph, checkpoint_interval=checkpoint_interval)
    assert L.hdpo_param_count(C.byref(desc)) == flat.size
    ws_bytes = L.hdpo_rollout_workspace_bytes(C.byref(desc))
    assert ws_bytes > 0, L.hdpo_last_error()
    ws = be.zeros(ws_bytes, np.uint8)
    if hasattr(be, "torch"):  # a dirty workspace: the kernels must not rely on zero-initialised scratch
        ws.view(be.torch.float32)[: (ws_bytes // 4)].fill_(float("nan"))
    B = bt.B
    p = be.ptr
    flat_h = be.put(flat)
    cost_b, report_b = be.zeros(B), be.zeros(B)
    reward_tb = be.zeros((T, B))
    totals = be.zeros(2, np.float64)
    fin_store = be.zeros(data["initial_inventories"].shape)
    fin_wh = be.zeros(data["initial_warehouse_inventories"].shape) if bt.W else None
    fin_ech = be.zeros(data["initial_echelon_inventories"].shape) if bt.E else None
    fin = K.State(p(fin_store), p(fin_wh), p(fin_ech))
    rc = L.hdpo_rollout_fwd(C.byref(desc), p(flat_h), p(dem_h), C.byref(bt.st), C.byref(bt.init), p(cost_b),
                            p(report_b), p(reward_tb), p(totals), C.byref(fin), p(ws), ws_bytes, be.stream)
    K.check(L, rc, "hdpo_rollout_fwd")
    be.sync()
    out = {"workspace_bytes": ws_bytes, "cost_b": be.get(cost_b), "report_b": be.get(report_b), "reward_tb": be.get(reward_tb),
           "totals": be.get(totals),
           "final": {"store": be.get(fin_store), "wh": be.get(fin_wh), "ech": be.get(fin_ech)}}
    if backward:
        if g_total is None:
            g_total = 1.0 / (B * T * bt.S)
        grad = be.put(np.full(flat.size, np.nan, np.float32))
        rc = L.hdpo_rollout_bwd(C.byref(desc), p(flat_h), p(dem_h), C.byref(bt.st), g_total, g_report, p(grad),
                                p(ws), ws_bytes, be.stream)
        K.check(L, rc, "hdpo_rollout_bwd")
        be.sync()
        out["grad_flat"] = be.get(grad)
        out["grad"] = unflatten(out["grad_flat"], names)
    return out


def step(be, meta, data, action, up=None, t=0):
    """hdpo_step_fwd (+ hdpo_step_bwd when `up` holds the upstream adjoints)."""
    L = be.lib
    pp = meta["problem_params"]
    bt = Batch(be, pp, data)
    p = be.ptr
    a = {k: be.put(v) for k, v in action.items()}
    act = K.Action(p(a["stores"]), p(a.get("warehouses")), p(a.get("echelons")))
    dem = be.put(data["demands"])
    Tt = data["demands"].shape[2]
    col = t + meta.get("period_shift", 0)
    dem_ptr = p(dem) + 4 * col
    nxt_store = be.zeros(data["initial_inventories"].shape)
    nxt_wh = be.zeros(data["initial_warehouse_inventories"].shape) if bt.W else None
    nxt_ech = be.zeros(data["initial_echelon_inventories"].shape) if bt.E else None
    nxt = K.State(p(nxt_store), p(nxt_wh), p(nxt_ech))
    reward = be.zeros(bt.B)
    rc = L.hdpo_step_fwd(C.byref(bt.pb), C.byref(bt.st), C.byref(bt.init), C.byref(act), dem_ptr, bt.S * Tt, Tt,
                         C.byref(nxt), p(reward), be.stream)
    K.check(L, rc, "hdpo_step_fwd")
    be.sync()
    out = {"reward": be.get(reward), "new": {"store": be.get(nxt_store), "wh": be.get(nxt_wh), "ech": be.get(nxt_ech)}}
    if up is not None:
        gn = {k: be.put(v) for k, v in up.items()}
        g_next = K.State(p(gn.get("store_inventories")), p(gn.get("warehouse_inventories")),
                         p(gn.get("echelon_inventories")))
        gc_store = be.zeros(data["initial_inventories"].shape)
        gc_wh = be.zeros(data["initial_warehouse_inventories"].shape) if bt.W else None
        gc_ech = be.zeros(data["initial_echelon_inventories"].shape) if bt.E else None
        ga = {k: be.zeros(v.shape) for k, v in action.items()}
        g_cur = K.State(p(gc_store), p(gc_wh), p(gc_ech))
        g_act = K.Action(p(ga["stores"]), p(ga.get("warehouses")), p(ga.get("echelons")))
        rc = L.hdpo_step_bwd(C.byref(bt.pb), C.byref(bt.st), C.byref(bt.init), C.byref(act), dem_ptr, bt.S * Tt, Tt,
                             C.byref(g_next), p(gn["reward"]), C.byref(g_cur), C.byref(g_act), be.stream)
        K.check(L, rc, "hdpo_step_bwd")
        be.sync()
        out["g_cur"] = {"store": be.get(gc_store), "wh": be.get(gc_wh), "ech": be.get(gc_ech)}
        out["g_act"] = {k: be.get(v) for k, v in ga.items()}
    return out
