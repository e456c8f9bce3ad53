"""GPU box: host time needed to ISSUE one forward+adjoint step vs the device time of the step (launch-bound check)."""
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
for _ in range(3):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
l0 = eng.lib.hdpo_kernel_launch_count()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    eng.forward(flat, data); eng.backward(g, 0.0, out=grad)
t1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
print(f"{name}: host issue {1e3 * (t1 - t0) / n:.2f} ms/step, device {e0.elapsed_time(e1) / n:.2f} ms/step, "
      f"{(eng.lib.hdpo_kernel_launch_count() - l0) // n} launches/step")
