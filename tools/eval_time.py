"""GPU box: forward-only long-horizon evaluation (Trainer.test shape: 32768 scenarios x 5000 periods, 3000 ignored,
discrete allocation) through the fused rollout, device-timed. SURVEY.md 8(f)-3."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from neural_inventory_control_b200 import engine as EN, workloads as WL

dev = torch.device("cuda", 0)
for name, B, T in (("one_store_lost", 32768, 5000), ("serial_system", 32768, 5000),
                   ("one_warehouse_lost_demand", 8192, 1000), ("one_warehouse_lost_demand_symmetry_aware", 8192, 1000)):
    pspec, pp, data, widths = WL.WORKLOADS[name](dev, B=B, T=T, seed=57)
    flat = WL.init_params(widths, torch.Generator(device=dev).manual_seed(0), dev)
    eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=int(0.6 * T), discrete_allocation=(name == "one_store_lost"),
                          precision="tf32x3", save_for_backward=False)
    for _ in range(2):
        eng.forward(flat, data)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 3
    for _ in range(n):
        tot = eng.forward(flat, data)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {B} scenarios x {T} periods forward-only: {ms:.1f} ms, {B * T / ms / 1e3:.1f} M scenario-periods/s, "
          f"workspace {eng.ws_bytes / 2**20:.0f} MiB, reported cost {float(tot[1]) / (B * (T - int(0.6 * T)) * pp['n_stores']):.3f}")
    del eng, data
    torch.cuda.empty_cache()
