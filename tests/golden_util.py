"""Helpers shared by the oracle tests (CPU) and the parity tests (GPU): load tests/golden/*.npz."""
import glob
import json
import os

import numpy as np

from oracle import hdpo_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rollout_cases():
    return sorted(os.path.basename(p)[len("rollout_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rollout_*.npz")))


def step_cases():
    return sorted(os.path.basename(p)[len("step_"):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "step_*.npz")))


def load(kind, name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    groups = {}
    for k in z.files:
        if k == "meta":
            continue
        head, rest = k.split("/", 1)
        groups.setdefault(head, {})[rest] = z[k]
    return meta, groups


def problem_from_meta(meta):
    pp = meta["problem_params"]
    return O.Problem(pp["n_stores"], pp["n_warehouses"], pp["n_extra_echelons"], bool(pp["lost_demand"]),
                     bool(pp["maximize_profit"]), int(meta.get("period_shift", 0)))


def policy_from_golden(meta, params, dtype=np.float32):
    nets = {}
    for module in meta["neurons_per_hidden_layer"]:
        net = O.mlp_from_state_dict(params, module, meta["inner_layer_activations"][module],
                                    meta["output_layer_activation"][module])
        nets[module] = net.astype(dtype)
    adj = meta["problem_params"].get("warehouse_store_adjacency")
    return O.Policy(meta["nn_name"], nets, meta["warehouse_upper_bound"],
                    None if adj is None else np.asarray(adj), meta.get("transshipment", False))


def cast(d, dtype):
    return {k: v.astype(dtype) for k, v in d.items()}


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
