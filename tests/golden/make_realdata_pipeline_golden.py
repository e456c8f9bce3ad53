#!/usr/bin/env python
"""Golden of the REAL-DATA dataset pipeline, produced by running the UNMODIFIED reference (data_handling.py:9-123,
162-168, 398-458) on two small fixture files committed next to this script:

    tests/golden/realdata_files/weekly_sales.pt       [24 samples, 3 stores, 60 weeks] float32 (synthetic "sales")
    tests/golden/realdata_files/dates_with_info.csv   60 weekly rows with a days_from_christmas column

`Scenario(...)` reads them (demand distribution 'real', time feature days_from_christmas), `DatasetCreator` splits the
result BY PERIOD into train / dev / test windows; the sha256 of every tensor of every split goes to
tests/golden/realdata_pipeline_hashes.json, which tests/test_real_data.py::test_real_data_pipeline_matches_reference
reproduces with this repo's data_handling on the same files. Run in the build container only:

    python tests/golden/make_realdata_pipeline_golden.py
"""
import copy
import hashlib
import json
import os
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HDPO_REFERENCE_ROOT", "/root/reference")
FILES = os.path.join(HERE, "realdata_files")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def make_fixture_files():
    os.makedirs(FILES, exist_ok=True)
    rng = np.random.RandomState(20240423)
    n, s, t = 24, 3, 60
    season = 1.0 + 0.4 * np.sin(2 * np.pi * np.arange(t) / 52.0)
    base = rng.gamma(4.0, 2.5, size=(n, s, 1))
    sales = np.maximum(base * season[None, None, :] * rng.lognormal(0.0, 0.35, size=(n, s, t)), 0.0).round(3)
    torch.save(torch.tensor(sales, dtype=torch.float32), os.path.join(FILES, "weekly_sales.pt"))
    with open(os.path.join(FILES, "dates_with_info.csv"), "w") as f:
        f.write("date,days_from_christmas\n")
        for w in range(t):
            f.write(f"week{w:03d},{(7 * w + 7) % 365}\n")


SETTING = {
    "seeds": {"underage_cost": 28, "holding_cost": 73, "mean": 33, "coef_of_var": 92, "lead_time": 41, "demand": 57,
              "initial_inventory": 4839, "warehouse": 10},
    "sample_data_params": {"split_by_period": True, "train_periods": "(0, 40)", "dev_periods": "(28, 52)",
                           "test_periods": "(36, 60)"},
    "problem_params": {"n_stores": 3, "n_warehouses": 2, "n_extra_echelons": 0, "lost_demand": True,
                       "maximize_profit": True, "warehouse_store_adjacency": [[1, 1, 0], [0, 1, 1]]},
    "observation_params": {"include_warehouse_inventory": True,
                           "include_static_features": {"holding_costs": True, "underage_costs": True, "lead_times": True,
                                                       "mean": False, "std": False},
                           "demand": {"past_periods": 8, "period_shift": 8},
                           "time_features_file": None, "time_features": ["days_from_christmas"]},
    "store_params": {
        "demand": {"distribution": "real", "file_location": None, "sample_across_stores": False, "expand": False,
                   "clip": False, "decimals": 3},
        "lead_time": {"sample_across_stores": False, "vary_across_samples": False, "expand": True,
                      "value": [[3, 1], [2, 2], [1, 3]]},
        "holding_cost": {"sample_across_stores": True, "vary_across_samples": False, "expand": False, "range": [0.7, 1.3]},
        "underage_cost": {"sample_across_stores": True, "vary_across_samples": True, "expand": False, "range": [6.3, 11.7]},
        "initial_inventory": {"sample": False, "inventory_periods": 4}},
    "warehouse_params": {"holding_cost": [0.3, 0.4], "lead_time": 3, "edge_cost": [0.5, 1.5]},
    "echelon_params": None,
}


def setting_with_paths():
    s = copy.deepcopy(SETTING)
    s["observation_params"]["time_features_file"] = os.path.join(FILES, "dates_with_info.csv")
    s["store_params"]["demand"]["file_location"] = os.path.join(FILES, "weekly_sales.pt")
    return s


def digest(t):
    return [list(t.shape), hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]]


def run(Scenario, DatasetCreator, n_samples=20, periods=60):
    s = setting_with_paths()
    obs = defaultdict(lambda: None, s["observation_params"])
    sc = Scenario(periods, s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"], n_samples,
                  obs, copy.deepcopy(s["seeds"]))
    windows = [s["sample_data_params"][k] for k in ("train_periods", "dev_periods", "test_periods")]
    parts = DatasetCreator().create_datasets(sc, split=True, by_period=True, periods_for_split=windows)
    out = {"full": {k: digest(v) for k, v in sc.get_data().items()}}
    for name, ds in zip(("train", "dev", "test"), parts):
        out[name] = {k: digest(v) for k, v in ds.data.items()}
        out[name]["__len__"] = len(ds)
    out["split_by"] = {k: sorted(v) for k, v in sc.define_how_to_split_data().items()} if hasattr(sc, "define_how_to_split_data") else None
    return out


if __name__ == "__main__":
    make_fixture_files()
    import trainer as ref_trainer  # the reference's star-import chain
    gold = run(ref_trainer.Scenario, ref_trainer.DatasetCreator)
    with open(os.path.join(HERE, "realdata_pipeline_hashes.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    for part, d in gold.items():
        if part != "split_by":
            print(part, {k: v[0] if isinstance(v, list) else v for k, v in d.items()})
