"""GPU box: does replaying one forward+adjoint step as a CUDA graph (all chunk streams captured) beat stream launches?"""
import sys, time
sys.path.insert(0, "/root/repo")
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
def step():
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
for _ in range(3):
    step()
torch.cuda.synchronize()
ref = grad.clone()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"stream launches: {timeit(step):.2f} ms/step")
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s, capture_error_mode="relaxed"):
        step()
torch.cuda.synchronize()
grad.zero_()
print(f"graph replay:    {timeit(gr.replay):.2f} ms/step")
print("gradient identical after replay:", bool(torch.equal(grad, ref)), float((grad - ref).abs().max()))
