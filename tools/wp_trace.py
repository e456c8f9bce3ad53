"""GPU box: per-role event timeline of the persistent wide sweeps (wide_persist.cu) for one pair.
Roles: 0 TMA producer, 1 MMA issuer, 2 first epilogue warp, 3 first head warp. Tags: 0x1LL reached task of layer LL,
0x2LL inputs ready / first operands landed / accumulators drained, 0x3LL loads issued / MMAs issued / tile staged,
0x4LL published."""
import ctypes as C
import os
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL

dev = torch.device("cuda", 0)
name = os.environ.get("WL", "one_warehouse_lost_demand")
T = int(os.environ.get("T", "6"))
pair_show = int(os.environ.get("PAIR", "0"))
pspec, pp, data, widths = WL.WORKLOADS[name](dev, seed=57, T=T)
B, S = data["demands"].shape[0], pp["n_stores"]
flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=0, precision="tf32x3")
grad = torch.zeros_like(flat)
g = 1.0 / (B * T * S)
lib = eng.lib
for _ in range(2):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
CAP = 2048
for phase in ("forward", "backward"):
    buf = torch.zeros(74 * 4 * CAP * 2, dtype=torch.int64, device=dev)
    lib.hdpo_debug_set_wp_trace(buf.data_ptr(), CAP)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if phase == "forward":
        eng.forward(flat, data)
    else:
        eng.backward(g, 0.0, out=grad)
    e1.record()
    torch.cuda.synchronize()
    lib.hdpo_debug_set_wp_trace(None, 0)
    h = buf.cpu().numpy().reshape(74, 4, CAP, 2)
    print(f"=== {phase}: event time {e0.elapsed_time(e1) * 1e3:.1f} us for T={T}")
    base = min(int(h[pair_show, r, 0, 1]) for r in range(4) if h[pair_show, r, 0, 1] > 0)
    ev = []
    for r in range(4):
        for i in range(CAP):
            tag, t = int(h[pair_show, r, i, 0]), int(h[pair_show, r, i, 1])
            if t == 0:
                break
            ev.append((t - base, r, tag))
    ev.sort()
    names = ["prod", "mma ", "epi ", "head"]
    last = {}
    lim = int(os.environ.get("LINES", "140"))
    for t, r, tag in ev[:lim]:
        dt = t - last.get(r, 0)
        last[r] = t
        print(f"{t / 1e3:9.2f} us  {names[r]}  tag {tag:#05x}   (+{dt / 1e3:.2f} us in role)")
    # per-role time between consecutive 'reached task' marks
    for r in range(4):
        ts = [t for t, rr, tag in ev if rr == r and (tag >> 8) == 1]
        if len(ts) > 2:
            d = np.diff(ts)
            print(f"role {names[r]}: {len(ts)} tasks, median {np.median(d) / 1e3:.2f} us between task starts, total {(ts[-1] - ts[0]) / 1e3:.1f} us")
