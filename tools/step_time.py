"""GPU box: device time of ONE generic simulator period (K3: hdpo_step_fwd + hdpo_step_bwd) at a many-warehouses shape.
HDPO_STEP_THREAD=1 selects the thread-per-scenario kernels (the round-1 form) for comparison."""
import ctypes as C
import os
import sys
sys.path.insert(0, "/root/repo")
import torch
from neural_inventory_control_b200 import _capi as K, _lib, spec
from neural_inventory_control_b200.engine import _ptr, current_stream_ptr

dev = torch.device("cuda", 0)
lib = _lib.load()
B, S, W, L, Lw = int(os.environ.get("B", 8192)), 50, 3, 6, 3
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.rand(*s, generator=g, device=dev)  # noqa: E731
pb = spec.problem(B, S, W, 0, L, Lw, 0, True, False, True)
store, wh = r(B, S, L) * 5, r(B, W, Lw) * 30
a_s, a_w = r(B, S, W) * 2, r(B, W, 1) * 9
lt = torch.randint(1, L + 1, (B, S, W), generator=g, device=dev).float()
statics = {"holding_costs": r(B, S), "underage_costs": r(B, S) * 9, "lead_times": lt,
           "warehouse_lead_times": torch.full((B, W), 3.0, device=dev), "warehouse_holding_costs": r(B, W),
           "warehouse_edge_costs": r(B, W), "echelon_lead_times": None, "echelon_holding_costs": None, "mean": None,
           "std": None}
keys = ("holding_costs", "underage_costs", "lead_times", "warehouse_lead_times", "warehouse_holding_costs",
        "warehouse_edge_costs", "echelon_lead_times", "echelon_holding_costs", "mean", "std")
st = K.Statics(*[_ptr(statics[k]) for k in keys])
dem = r(B, S, 4) * 8
n_store, n_wh, reward = torch.empty_like(store), torch.empty_like(wh), torch.empty(B, device=dev)
cur, nxt = K.State(_ptr(store), _ptr(wh), None), K.State(_ptr(n_store), _ptr(n_wh), None)
act = K.Action(_ptr(a_s), _ptr(a_w), None)
gs, gw, gr = torch.randn_like(store), torch.randn_like(wh), r(B)
gn = K.State(_ptr(gs), _ptr(gw), None)
gc_s, gc_w, ga_s, ga_w = torch.empty_like(store), torch.empty_like(wh), torch.empty_like(a_s), torch.empty_like(a_w)
g_cur, g_act = K.State(_ptr(gc_s), _ptr(gc_w), None), K.Action(_ptr(ga_s), _ptr(ga_w), None)
sp = current_stream_ptr(dev)


def fwd():
    K.check(lib, lib.hdpo_step_fwd(C.byref(pb), C.byref(st), C.byref(cur), C.byref(act), dem.data_ptr(), dem.stride(0),
                                   dem.stride(1), C.byref(nxt), reward.data_ptr(), sp), "fwd")


def bwd():
    K.check(lib, lib.hdpo_step_bwd(C.byref(pb), C.byref(st), C.byref(cur), C.byref(act), dem.data_ptr(), dem.stride(0),
                                   dem.stride(1), C.byref(gn), gr.data_ptr(), C.byref(g_cur), C.byref(g_act), sp), "bwd")


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


# plain torch restatement of the order adjoints (environment.py:391-434 + the SURVEY 8a recurrences) for this shape
raw_w = wh[:, :, 0] - a_s.sum(dim=1)
g_raw_w = gr[:, None] * statics["warehouse_holding_costs"] * (raw_w >= 0).float() + gw[:, :, 0]
ga_ref = torch.gather(gs, 2, (lt - 1).long()) * (a_s != 0).float() - g_raw_w[:, None, :]
bwd()
torch.cuda.synchronize()
err = float((ga_s - ga_ref).abs().max())
form = "thread-per-scenario" if os.environ.get("HDPO_STEP_THREAD") == "1" else "warp-per-scenario"
print(f"K3 {form}: B={B} S={S} W={W} L={L}: fwd (incl. stray pass) {timeit(fwd):.1f} us, bwd {timeit(bwd):.1f} us; "
      f"reward checksum {float(reward.double().sum()):.6e} grad checksum {float(ga_s.double().sum()):.6e}, "
      f"max |g_act.stores - torch restatement| {err:.2e}")
