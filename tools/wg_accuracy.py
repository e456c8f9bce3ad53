"""GPU box: how the K-slice length of the weight-gradient GEMMs (HDPO_WG_KPS) changes the full-batch gradient.
Writes the flat gradient of one step (8192 x 50, one_warehouse_lost_demand, tf32x3) to gpurun_out/wg_grad_<kps>.npy;
`compare` prints rel-L2 distances between the runs and against the fp32 (SIMT, true fp32 accumulation) path."""
import os
import sys
sys.path.insert(0, "/root/repo")
import numpy as np

if len(sys.argv) > 1 and sys.argv[1] == "compare":
    tags = sys.argv[2:]
    g = {t: np.load(f"gpurun_out/wg_grad_{t}.npy").astype(np.float64) for t in tags}
    ref = g[tags[0]]
    for t in tags[1:]:
        d = np.linalg.norm(g[t] - ref) / np.linalg.norm(ref)
        print(f"rel-L2 |grad[{t}] - grad[{tags[0]}]| = {d:.3e}   max rel elem {np.abs(g[t] - ref).max() / np.abs(ref).max():.3e}")
    sys.exit(0)

import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL
dev = torch.device("cuda", 0)
tag = sys.argv[1]
precision = "fp32" if tag == "fp32" else "tf32x3"
pspec, pp, data, widths = WL.WORKLOADS["one_warehouse_lost_demand"](dev, seed=57, T=50)
B, S, T = data["demands"].shape[0], pp["n_stores"], 50
flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=30, precision=precision)
grad = torch.zeros_like(flat)
eng.forward(flat, data)
eng.backward(1.0 / (B * T * S), 0.0, out=grad)
torch.cuda.synchronize()
np.save(f"gpurun_out/wg_grad_{tag}.npy", grad.cpu().numpy())
print(tag, "done", float(grad.double().norm()))
