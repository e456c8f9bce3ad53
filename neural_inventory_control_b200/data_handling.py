"""Scenario / dataset construction with the reference's interface (data_handling.py:3-458).

Host-side, one-off work; kept as thin Python so `main_run.py` and user notebooks run unchanged. The numpy
legacy-RNG call order is part of the interface (same seeds => same demand traces / costs / inventories as the
reference; pinned by tests/golden/scenario_hashes.json): each generator below re-seeds `np.random` exactly where
the reference does (data_handling.py:183-185, 207-209, 231-235, 244, 295).
"""
import copy
from collections import defaultdict

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset


class Scenario:
    """One problem instance: sampled primitives, demand traces and initial inventories (data_handling.py:3-52)."""

    def __init__(self, periods, problem_params, store_params, warehouse_params, echelon_params, num_samples,
                 observation_params, seeds=None):
        self.problem_params = problem_params
        self.store_params = store_params
        self.warehouse_params = warehouse_params
        self.echelon_params = echelon_params
        self.num_samples = num_samples
        self.periods = periods
        self.observation_params = observation_params
        self.seeds = seeds
        n_stores = problem_params["n_stores"]

        # order matters: every draw below re-seeds numpy's legacy global generator
        self.demands = self.generate_demand_samples(problem_params, store_params, store_params["demand"], seeds)
        self.underage_costs = self._per_store(store_params["underage_cost"], seeds["underage_cost"], n_stores, False)
        self.holding_costs = self._per_store(store_params["holding_cost"], seeds["holding_cost"], n_stores, False)
        self.lead_times = self.generate_lead_times(problem_params, store_params["lead_time"], seeds["lead_time"])
        self.means, self.stds = self._static_demand_moments(observation_params, store_params)
        self.initial_inventories = self.generate_initial_inventories(problem_params, store_params, self.demands,
                                                                     self.lead_times, seeds["initial_inventory"])

        self.initial_warehouse_inventories = self._zero_pipeline(warehouse_params, problem_params["n_warehouses"])
        self.warehouse_lead_times = self._per_warehouse(warehouse_params, "lead_time")
        self.warehouse_holding_costs = self._per_warehouse(warehouse_params, "holding_cost")
        has_edge = bool(warehouse_params) and "edge_cost" in warehouse_params
        self.warehouse_edge_costs = self._per_warehouse(warehouse_params, "edge_cost") if has_edge else None

        self.initial_echelon_inventories = self._zero_pipeline(echelon_params, None)
        self.echelon_lead_times = self._per_echelon(echelon_params, "lead_time")
        self.echelon_holding_costs = self._per_echelon(echelon_params, "holding_cost")

        self.time_features, self.sample_features = self._read_feature_files(observation_params, n_stores)
        self.split_by = self.define_how_to_split_data()

    # ------------------------------------------------------------------ public API
    def get_data(self):
        """dict of float32 tensors, None entries dropped (data_handling.py:54-81); lead times are float-encoded."""
        fields = {
            "demands": self.demands, "underage_costs": self.underage_costs, "holding_costs": self.holding_costs,
            "lead_times": self.lead_times, "mean": self.means, "std": self.stds,
            "initial_inventories": self.initial_inventories,
            "initial_warehouse_inventories": self.initial_warehouse_inventories,
            "warehouse_lead_times": self.warehouse_lead_times,
            "warehouse_holding_costs": self.warehouse_holding_costs,
            "warehouse_edge_costs": self.warehouse_edge_costs,
            "initial_echelon_inventories": self.initial_echelon_inventories,
            "echelon_holding_costs": self.echelon_holding_costs, "echelon_lead_times": self.echelon_lead_times,
        }
        fields.update(self.time_features)
        fields.update(self.sample_features)
        return {k: v.float() for k, v in fields.items() if v is not None}

    def define_how_to_split_data(self):
        """Which keys split by sample index and which by period (data_handling.py:83-123)."""
        by_sample = ["underage_costs", "holding_costs", "lead_times", "initial_inventories"]
        by_period = []
        if self.problem_params["n_warehouses"] > 0:
            by_sample += ["initial_warehouse_inventories", "warehouse_lead_times", "warehouse_holding_costs",
                          "warehouse_edge_costs"]
        if self.problem_params["n_extra_echelons"] > 0:
            by_sample += ["initial_echelon_inventories", "echelon_holding_costs", "echelon_lead_times"]
        (by_period if self.store_params["demand"]["distribution"] == "real" else by_sample).append("demands")
        static = self.observation_params["include_static_features"]
        by_sample += [k for k in ("mean", "std") if static.get(k)]
        by_period += list(self.time_features)
        by_sample += list(self.sample_features)
        return {"sample_index": by_sample, "period": by_period}

    # ------------------------------------------------------------------ demand
    def generate_demand_samples(self, problem_params, store_params, demand_params, seeds):
        if demand_params["sample_across_stores"]:  # writes mean/std into the caller's dict, like the reference
            demand_params.update(self.sample_normal_mean_and_std(problem_params, demand_params, seeds))
        self.adjust_seeds_for_consistency(problem_params, store_params, seeds)
        draw = {"normal": self.generate_normal_demand, "poisson": self.generate_poisson_demand,
                "real": self.read_real_demand_data}[demand_params["distribution"]]
        demand = draw(problem_params, demand_params, seeds["demand"])
        if demand_params["clip"]:
            demand = np.clip(demand, 0, None)
        return torch.tensor(demand)

    def adjust_seeds_for_consistency(self, problem_params, store_params, seeds):
        """One-store synthetic settings shift the demand seed IN PLACE by int(lead + 10*underage)
        (data_handling.py:150-160)."""
        single = problem_params["n_warehouses"] == 0 and problem_params["n_stores"] == 1
        if single and store_params["demand"]["distribution"] != "real":
            try:
                seeds["demand"] += int(store_params["lead_time"]["value"] + 10 * store_params["underage_cost"]["value"])
            except Exception as e:  # noqa: BLE001 - same best-effort behaviour as the reference
                print(f"Error: {e}")

    def read_real_demand_data(self, problem_params, demand_params, seed):
        return torch.load(demand_params["file_location"])[: self.num_samples]

    def generate_normal_demand(self, problem_params, demand_params, seed):
        if seed is not None:
            np.random.seed(seed)
        n, T = self.num_samples, self.periods
        if problem_params["n_stores"] == 1:
            return np.random.normal(demand_params["mean"], demand_params["std"], size=(n, 1, T))
        std = np.asarray(demand_params["std"], dtype=float)
        rho = demand_params["correlation"]
        # element [j][i] = (rho*std_i)*std_j off the diagonal, std_i*std_j on it - same rounding as the reference
        cov = (rho * std)[None, :] * std[:, None]
        np.fill_diagonal(cov, std * std)
        draws = np.random.multivariate_normal(demand_params["mean"], cov=cov.tolist(), size=(n, T))
        return np.transpose(draws, (0, 2, 1))

    def generate_poisson_demand(self, problem_params, demand_params, seed):
        if seed is not None:
            np.random.seed(seed)
        return np.random.poisson(demand_params["mean"], size=(self.num_samples, problem_params["n_stores"], self.periods))

    def sample_normal_mean_and_std(self, problem_params, demand_params, seeds):
        n_stores = problem_params["n_stores"]
        np.random.seed(seeds["mean"])
        means = np.random.uniform(*demand_params["mean_range"][:2], n_stores).round(3)
        np.random.seed(seeds["coef_of_var"])
        cv = np.random.uniform(*demand_params["coef_of_var_range"][:2], n_stores)
        return {"mean": means, "std": (means * cv).round(3)}

    # ------------------------------------------------------------------ costs / lead times / inventories
    def _per_store(self, params, seed, n_stores, discrete):
        """data_handling.py:239-271 (`generate_data_for_samples_and_stores`)."""
        np.random.seed(seed)
        p = defaultdict(lambda: False, copy.deepcopy(params))
        draw = np.random.randint if discrete else np.random.uniform
        if p["file_location"]:
            p["value"] = torch.load(p["file_location"])[: self.num_samples]
        if p["sample_across_stores"]:
            return torch.tensor(draw(*p["range"], n_stores)).expand(self.num_samples, -1)
        if p["vary_across_samples"]:
            return torch.tensor(draw(*p["range"], self.num_samples)).unsqueeze(1).expand(-1, n_stores)
        if p["expand"]:
            value = torch.tensor(p["value"])
            if value.dim() == 2:  # [n_stores, n_warehouses] warehouse-to-store lead-time matrix
                return value.unsqueeze(0).expand(self.num_samples, -1, -1)
            return value.expand(self.num_samples, n_stores)
        return torch.tensor(p["value"])

    # the reference's name for the same helper (kept for API compatibility)
    def generate_data_for_samples_and_stores(self, problem_params, cost_params, seed, discrete=False):
        return self._per_store(cost_params, seed, problem_params["n_stores"], discrete)

    def generate_lead_times(self, problem_params, lead_time_params, seed):
        """Always 3-D [num_samples, n_stores, max(n_warehouses,1)] int64 (data_handling.py:273-288)."""
        raw = self._per_store(lead_time_params, seed, problem_params["n_stores"], True)
        if raw.dim() == 2:
            width = max(problem_params.get("n_warehouses", 0), 1)
            raw = raw.unsqueeze(2).expand(-1, -1, width)
        return raw.to(torch.int64)

    def generate_initial_inventories(self, problem_params, store_params, demands, lead_times, seed):
        np.random.seed(seed)
        spec = store_params["initial_inventory"]
        n_stores = problem_params["n_stores"]
        if not spec["sample"]:
            return torch.zeros(self.num_samples, n_stores, spec["inventory_periods"])
        mean_demand = demands.float().mean(dim=2).mean(dim=0)
        width = max(spec["inventory_periods"], lead_times.max().item())
        mult = np.random.uniform(*spec["range_mult"], size=(self.num_samples, n_stores, width))
        return mean_demand[None, :, None] * mult

    def _zero_pipeline(self, params, n_nodes):
        if params is None:
            return None
        lead = params["lead_time"]
        if n_nodes is None:  # echelons: one entry per echelon
            return torch.zeros(self.num_samples, len(lead), max(lead))
        return torch.zeros(self.num_samples, n_nodes, max(lead) if isinstance(lead, list) else lead)

    def _per_warehouse(self, params, key):
        if params is None:
            return None
        n = self.problem_params["n_warehouses"]
        value = params[key]
        if isinstance(value, list):
            if len(value) != n:
                raise ValueError(f"warehouse_params['{key}'] list length {len(value)} doesn't match n_warehouses {n}")
            return torch.tensor(value).unsqueeze(0).expand(self.num_samples, -1)
        return torch.tensor([value]).expand(self.num_samples, n)

    def _per_echelon(self, params, key):
        if params is None:
            return None
        return torch.tensor(params[key]).unsqueeze(0).expand(self.num_samples, -1)

    def _static_demand_moments(self, observation_params, store_params):
        out = []
        for k in ("mean", "std"):
            if observation_params["include_static_features"].get(k):
                out.append(torch.tensor(store_params["demand"][k]).unsqueeze(0).expand(self.num_samples, -1))
            else:
                out.append(None)
        return out

    def _read_feature_files(self, observation_params, n_stores):
        found = {"time_features": {}, "sample_features": {}}
        for kind in found:
            path = observation_params.get(f"{kind}_file")
            if not (observation_params.get(kind) and path):
                continue
            table = pd.read_csv(path)
            for name in observation_params[kind]:
                col = torch.tensor(table[name].values)
                if kind == "time_features":
                    found[kind][name] = col.unsqueeze(0).unsqueeze(0).expand(self.num_samples, n_stores, -1)
                else:  # one value per sample (one-store settings only)
                    found[kind][name] = col.unsqueeze(1).expand(-1, n_stores)
        return found["time_features"], found["sample_features"]


class MyDataset(Dataset):
    """dict-of-tensors dataset (data_handling.py:385-395)."""

    def __init__(self, num_samples, data):
        self.data = data
        self.num_samples = num_samples

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        return {k: v[idx] for k, v in self.data.items()}


class DatasetCreator:
    """Train/dev split by sample index or by period (data_handling.py:398-458)."""

    def create_datasets(self, scenario, split=True, by_period=False, by_sample_indexes=False, periods_for_split=None,
                        sample_index_for_split=None):
        if not split:
            return self.create_single_dataset(scenario.get_data())
        if by_period:
            return [self.create_single_dataset(d) for d in self.split_by_period(scenario, periods_for_split)]
        if by_sample_indexes:
            train, dev = self.split_by_sample_index(scenario, sample_index_for_split)
            return self.create_single_dataset(train), self.create_single_dataset(dev)
        raise NotImplementedError

    def split_by_sample_index(self, scenario, sample_index_for_split):
        """The FIRST rows are the dev set, so the dev set does not depend on the train-set size."""
        data = scenario.get_data()
        n = sample_index_for_split
        return {k: v[n:] for k, v in data.items()}, {k: v[:n] for k, v in data.items()}

    def split_by_period(self, scenario, periods_for_split):
        data = scenario.get_data()
        shared = {k: data[k] for k in scenario.split_by["sample_index"] if k in data}
        out = []
        for text in periods_for_split:  # e.g. "(16, 136)" -> slice(16, 136)
            window = slice(*(int(x) for x in text.strip("() ").split(",")))
            part = copy.deepcopy(shared)
            for k in scenario.split_by["period"]:
                if k in data:
                    part[k] = data[k][:, :, window]
            out.append(part)
        return out

    def create_single_dataset(self, data):
        return MyDataset(len(data["initial_inventories"]), data)
