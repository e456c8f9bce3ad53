"""GPU box: run exactly ONE forward + adjoint step of a bench workload inside a cudaProfilerStart/Stop window
(3 warm-up steps before it), for `ncu --profile-from-start off ...` launch lists and `--set full` captures.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_one_step.py <workload> [precision] [batch]
"""
import sys
sys.path.insert(0, "/root/repo")
import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL

dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else "one_warehouse_lost_demand"
pspec, pp, data, widths = WL.WORKLOADS[name](dev, seed=57, T=50, **({"B": int(sys.argv[3])} if len(sys.argv) > 3 else {}))
B, S, T = data["demands"].shape[0], pp["n_stores"], 50
flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
small = pspec.arch in ("vanilla_one_store", "vanilla_serial")
precision = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=30, precision=precision)
grad = torch.zeros_like(flat)
g = 1.0 / (B * T * S)
for _ in range(3):
    eng.forward(flat, data)
    eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(flat, data)
eng.backward(g, 0.0, out=grad)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step of", name, precision, "B =", B)
