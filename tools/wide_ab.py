"""GPU box: forward / adjoint / step time of one wide-path workload under the current HDPO_* environment knobs."""
import os
import sys
sys.path.insert(0, "/root/repo")
import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL
dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else "one_warehouse_lost_demand"
kw = {}
if os.environ.get("HDPO_AB_BATCH"):
    kw = {"B": int(os.environ["HDPO_AB_BATCH"])}
if name.endswith("_8192"):
    name, kw = name[:-5], {"B": 8192}
pspec, pp, data, widths = WL.WORKLOADS[name](dev, seed=57, T=50, **kw)
B, S, T = data["demands"].shape[0], pp["n_stores"], 50
flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=30, precision=os.environ.get("HDPO_AB_PRECISION", "tf32x3"))
grad = torch.zeros_like(flat)
g = 1.0 / (B * T * S)
for _ in range(3):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
N = 10
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * N)]
torch.cuda.synchronize()
for i in range(N):
    ev[3 * i].record(); eng.forward(flat, data); ev[3 * i + 1].record(); eng.backward(g, 0.0, out=grad); ev[3 * i + 2].record()
torch.cuda.synchronize()
f = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(N)) / N
b = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(N)) / N
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
e1.record(); torch.cuda.synchronize()
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("HDPO_"))
print(f"{name} [{knobs}] fwd {f:.2f} bwd {b:.2f} step {e0.elapsed_time(e1) / N:.2f} ms")
