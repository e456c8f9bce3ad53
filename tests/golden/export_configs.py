#!/usr/bin/env python
"""Re-emit the reference's YAML settings in this repo's normalised form (config_files/ is the schema/API).

The YAML schema and the shipped values are part of the public surface `main_run.py <mode> <setting> <policy>`
resolves by name (main_run.py:24-40 in the reference). This script loads the in-scope reference files, and
writes them back sorted, comment-free and flow-styled, so that values stay identical while the text is ours.
Run in the build container only. Also writes tests/golden/scenario_hashes.json: sha256 of every tensor of
`Scenario.get_data()` produced by the UNMODIFIED reference for each setting (64 samples x 60 periods), which
pins this repo's data_handling.Scenario (same seeds => bit-identical data).
"""
import copy
import hashlib
import json
import os
import sys
from collections import defaultdict

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HDPO_REFERENCE_ROOT", "/root/reference")
SETTINGS = ["one_store_lost", "one_store_backlogged", "serial_system", "one_warehouse_lost_demand",
            "many_warehouses_lost_demand", "transshipment_backlogged"]
POLICIES = ["vanilla_one_store", "vanilla_serial", "vanilla_warehouse", "vanilla_transshipment", "base_stock",
            "capped_base_stock", "echelon_stock"]


def emit(kind, name):
    with open(f"{REF}/config_files/{kind}/{name}.yml") as f:
        cfg = yaml.safe_load(f)
    out_dir = os.path.join(ROOT, "config_files", kind)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"{name}.yml"), "w") as f:
        f.write(f"# {kind}/{name}: values as shipped by the reference; schema documented in DESIGN.md\n")
        yaml.safe_dump(cfg, f, sort_keys=True, default_flow_style=None, width=110)
    return cfg


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
    sys.path.insert(0, REF)
    import trainer as ref  # the reference
    hashes = {}
    for name in SETTINGS:
        cfg = emit("settings", name)
        s = copy.deepcopy(cfg)
        obs = defaultdict(lambda: None, s["observation_params"])
        sc = ref.Scenario(60, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], 64,
                          obs, s["seeds"])
        hashes[name] = {k: [list(v.shape), hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest()[:16]]
                        for k, v in sc.get_data().items()}
        hashes[name]["__seeds_after__"] = s["seeds"]
    for name in POLICIES:
        emit("policies_and_hyperparams", name)
    with open(os.path.join(HERE, "scenario_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("wrote", len(SETTINGS), "settings,", len(POLICIES), "policies, scenario_hashes.json")


if __name__ == "__main__":
    main()
