"""Debug helper (GPU box): per-tensor gradient error of the wide rollout in fp32 / tf32x3 mode vs the float64 oracle."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import abi_driver as D, golden_util as G
from oracle import hdpo_oracle as O
be = D.CudaBackend()
meta, g = G.load("rollout", "one_warehouse_s50")
rng = np.random.RandomState(0)
widths = [153, 512, 512, 51] if len(sys.argv) < 2 else [int(x) for x in sys.argv[1].split(",")]
B = 200 if len(sys.argv) < 3 else int(sys.argv[2])
T = 5 if len(sys.argv) < 4 else int(sys.argv[3])
params = {}
for i in range(len(widths) - 1):
    k = 1 / np.sqrt(widths[i])
    params[f"net.master.{2 * i}.weight"] = rng.uniform(-k, k, (widths[i + 1], widths[i])).astype(np.float32)
    params[f"net.master.{2 * i}.bias"] = rng.uniform(-k, k, (widths[i + 1],)).astype(np.float32)
data = {k: np.concatenate([v] * 13, 0)[:B] for k, v in g["data"].items()}
pb = G.problem_from_meta(meta)
meta = dict(meta); meta["neurons_per_hidden_layer"] = {"master": widths[1:-1]}
pol = G.policy_from_golden(meta, params, np.float64)
fwd, grads = O.rollout_grad(pol, pb, G.cast(data, np.float64), T)
flat = O.flatten_grads(pol, grads)
for prec in ("fp32", "tf32x3"):
    out = D.rollout(be, meta, params, data, T=T, ignore=1, precision=prec)
    print(prec, "cost err", np.abs(out["cost_b"] / fwd["reward_tb"].sum(0) - 1).max())
    for k in sorted(flat):
        print(f"   {k:24s} rel-L2 err {G.rel_l2(out['grad'][k], flat[k]):.3e}   |g| {np.linalg.norm(flat[k]):.3e}")
