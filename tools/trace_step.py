"""GPU box: per-CTA timeline (globaltimer) of one forward + adjoint step of the wide rollout: who occupies the SMs when.
Prints SM-time by kernel kind, the number of busy SMs over time and the dependency chain of one chunk-period."""
import ctypes as C
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL

dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else "one_warehouse_lost_demand"
pspec, pp, data, widths = WL.WORKLOADS[name](dev, seed=57, T=50)
B, S, T = data["demands"].shape[0], pp["n_stores"], 50
flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=30, precision="tf32x3")
grad = torch.zeros_like(flat)
g = 1.0 / (B * T * S)
lib = eng.lib
lib.hdpo_debug_set_trace.argtypes = [C.c_void_p, C.c_int64]
for _ in range(3):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
CAP = 400000
EPI = {0: "fwd_hidden", 1: "fwd_out", 2: "dgrad_hidden", 3: "dgrad_accum", 4: "store"}

def kind(tag):
    if (tag >> 8) == 0xF:
        return "head"
    low = tag & 0xFF
    k = EPI[(low >> 4) & 7]
    if low & 8:
        k = "wgrad"
    return k

for phase in ("forward", "backward"):
    buf = torch.zeros(4 + 4 * CAP, dtype=torch.int64, device=dev)
    lib.hdpo_debug_set_trace(buf.data_ptr(), CAP)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if phase == "forward":
        eng.forward(flat, data)
    else:
        eng.backward(g, 0.0, out=grad)
    e1.record()
    torch.cuda.synchronize()
    lib.hdpo_debug_set_trace(None, 0)
    h = buf.cpu().numpy()
    n = int(h[0] & 0xFFFFFFFF)
    rec = h[4:4 + 4 * n].reshape(n, 4)
    t0, t1 = rec[:, 0].astype(np.float64), rec[:, 1].astype(np.float64)
    smid = rec[:, 2] & 0xFFFF
    tag = rec[:, 2] >> 16
    base = t0.min()
    t0 -= base; t1 -= base
    span = t1.max()
    print(f"=== {phase}: {n} CTA records, traced span {span / 1e3:.1f} us, event time {e0.elapsed_time(e1) * 1e3:.1f} us, SMs seen {len(set(smid.tolist()))}")
    kinds = np.array([kind(int(x)) for x in tag])
    for k in sorted(set(kinds)):
        m = kinds == k
        d = (t1 - t0)[m]
        extra = ""
        if k != "head":
            aux = (rec[:, 3] >> 32)[m]
            ph = [((aux >> sh) & 1023).astype(np.float64).mean() * 32 / 1e3 for sh in (0, 10, 20)]
            extra = f"  | accumulators ready {ph[0]:5.2f}, tile staged {ph[1]:5.2f}, stores read {ph[2]:5.2f} us after start"
        print(f"   {k:13s} CTAs {m.sum():7d}  mean {d.mean() / 1e3:7.2f} us  SM-time {d.sum() / 1e6:8.2f} ms  = {d.sum() / span / 148 * 100:5.1f}% of 148 SMs x span{extra}")
    hm = kinds == "head"
    if phase == "forward" and hm.any():
        staged = (rec[:, 3] >> 32)[hm].astype(np.float64)
        print(f"   head_fwd: time until the staged rows landed: mean {staged.mean() / 1e3:.2f} us (of {(t1 - t0)[hm].mean() / 1e3:.2f} us)")
    # busy SM count over time, GEMM-class kernels only (1 CTA per SM)
    gm = kinds != "head"
    ev = np.concatenate([np.stack([t0[gm], np.ones(gm.sum())], 1), np.stack([t1[gm], -np.ones(gm.sum())], 1)])
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    active = np.cumsum(ev[:, 1])
    dt = np.diff(ev[:, 0], append=ev[-1, 0])
    hist = np.zeros(6)
    edges = [0, 1, 32, 64, 100, 140, 10 ** 9]
    for i in range(6):
        m = (active >= edges[i]) & (active < edges[i + 1])
        hist[i] = dt[m].sum()
    print("   time share by number of resident GEMM CTAs: " + ", ".join(
        f"[{edges[i]},{min(edges[i + 1], 149)}) {hist[i] / span * 100:.1f}%" for i in range(6)),
        f"; mean {(active * dt).sum() / span:.1f}")
    # the chain of chunk 0 in the middle of the sweep: first CTA start / last CTA end per kernel launch
    c0 = ((tag >> 8) & 0xFF) == 0
    c0 |= (tag == 0xF00)
    order = np.argsort(t0)
    mid = span * 0.5
    rows = []
    last_kind, s, e, cnt = None, 0, 0, 0
    for i in order:
        if not c0[i] or t0[i] < mid:
            continue
        k = kinds[i]
        if k != last_kind or t0[i] > e + 1:
            if last_kind is not None:
                rows.append((last_kind, s, e, cnt))
            last_kind, s, e, cnt = k, t0[i], t1[i], 0
        e = max(e, t1[i]); cnt += 1
        if len(rows) >= 12:
            break
    print("   chunk-0 chain mid-sweep (kernel: start, duration, gap to previous end; us):")
    prev = None
    for k, s, e, cnt in rows:
        print(f"      {k:13s} {cnt:4d} CTAs  start {(s - mid) / 1e3:8.2f}  dur {(e - s) / 1e3:6.2f}  gap {((s - prev) / 1e3) if prev is not None else 0:6.2f}")
        prev = e
