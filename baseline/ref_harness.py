"""Run the UNMODIFIED reference (its own Trainer.simulate_batch + backward, trainer.py:163-173) on a synthetic workload
of this repo - the timed baseline of `bench.py --impl reference` and of the `reference_cuda` entry of the bench line.

The reference modules come from baseline/_ref (byte-compiled by baseline/build_ref.py in the build container) or,
where it exists, from /root/reference. `gymnasium` / `matplotlib` are absent from the image: the two inert import stubs
of oracle/refstubs are put on the path (bench.py's reference arm is one of the places allowed to use oracle/).
Nothing of this repo's engine is on that path: the reference's own policy classes, Simulator and Trainer do the work;
only the synthetic tensors and the initial weights (so that both arms run the same numbers) are handed over.
"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


REF_MODULES = ("shared_imports", "quantile_forecaster", "data_handling", "neural_networks", "environment",
               "loss_functions", "trainer")  # dependency order of the reference's star-import chain


def locate():
    built = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(built, "trainer.bin")):
        return built
    src = os.environ.get("HDPO_REFERENCE_ROOT", "/root/reference")
    if os.path.exists(os.path.join(src, "trainer.py")):
        return src
    return None


_REF = None


def load_reference():
    """Import the reference's modules (byte-compiled .bin files of baseline/_ref, else the sources of /root/reference)
    under their own names for the duration of the import, then restore this repo's same-named root shims. Returns
    (the reference's `trainer` module - it star-imports everything else -, where it came from)."""
    global _REF
    if _REF is not None:
        return _REF
    where = locate()
    if where is None:
        raise ImportError("the reference is neither in baseline/_ref nor in /root/reference")
    import importlib.machinery
    import importlib.util
    sys.dont_write_bytecode = True
    saved = {m: sys.modules.pop(m) for m in REF_MODULES if m in sys.modules}
    stubs = os.path.join(ROOT, "oracle", "refstubs")  # inert `gymnasium` / `matplotlib` (absent from the image)
    sys.path.insert(0, stubs)
    loaded = {}
    try:
        for name in REF_MODULES:
            path_bin, path_py = os.path.join(where, name + ".bin"), os.path.join(where, name + ".py")
            if os.path.exists(path_bin):
                loader = importlib.machinery.SourcelessFileLoader(name, path_bin)
            else:
                loader = importlib.machinery.SourceFileLoader(name, path_py)
            spec = importlib.util.spec_from_loader(name, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            loader.exec_module(mod)
            loaded[name] = mod
    finally:
        sys.path.remove(stubs)
        for name in REF_MODULES:
            sys.modules.pop(name, None)
        sys.modules.update(saved)
    _REF = (loaded["trainer"], where)
    return _REF


_NN = {
    "vanilla_warehouse": lambda widths: {
        "name": "vanilla_warehouse", "inner_layer_activations": {"master": "elu"},
        "output_layer_activation": {"master": None}, "neurons_per_hidden_layer": {"master": list(widths[1:-1])},
        "output_sizes": {"master": None}, "initial_bias": None, "warehouse_upper_bound_mult": 4},
    "vanilla_one_store": lambda widths: {
        "name": "vanilla_one_store", "inner_layer_activations": {"master": "elu"},
        "output_layer_activation": {"master": None}, "neurons_per_hidden_layer": {"master": list(widths[1:-1])},
        "output_sizes": {"master": 1}, "initial_bias": None},
    "vanilla_serial": lambda widths: {
        "name": "vanilla_serial", "inner_layer_activations": {"master": "elu"},
        "output_layer_activation": {"master": None}, "neurons_per_hidden_layer": {"master": list(widths[1:-1])},
        "output_sizes": {"master": widths[-1]}, "initial_bias": None},
}


class ReferenceRun:
    """The reference's model / Simulator / Trainer on one synthetic batch (same tensors and weights as our arm)."""

    def __init__(self, workload, B, T, device="cpu", seed=57):
        import torch
        from collections import defaultdict
        from neural_inventory_control_b200 import workloads as WL
        self.torch = torch
        ref, self.where = load_reference()
        self.ref = ref
        self.device = device
        pspec, pp, data, widths = WL.WORKLOADS[workload]("cpu", B=B, T=T, seed=seed)
        if pspec.arch not in _NN:
            raise NotImplementedError(f"the reference snapshot has no '{pspec.arch}' policy (SURVEY.md section 0)")
        self.B, self.T, self.S = B, T, pp["n_stores"]
        self.problem_params = dict(pp)
        self.data = {k: v.to(device) for k, v in data.items()}
        self.obs_params = defaultdict(lambda: None, {
            "include_warehouse_inventory": pp["n_warehouses"] > 0,
            "include_static_features": {"holding_costs": True, "underage_costs": True, "lead_times": True},
            "demand": {"past_periods": 0, "period_shift": 0}})
        mean = data["mean"][0].tolist() if "mean" in data else [5.0]
        scenario = types.SimpleNamespace(problem_params=self.problem_params, store_params={"demand": {"mean": mean}})
        nn_params = _NN[pspec.arch](widths)
        torch.manual_seed(0)
        self.model = ref.NeuralNetworkCreator().create_neural_network(scenario, nn_params, device=device)
        if pspec.arch == "vanilla_serial":
            self.model.warehouse_upper_bound = torch.tensor([pspec.warehouse_upper_bound]).float().to(device)
        self.sim = ref.Simulator(device=device)
        self.trainer = ref.Trainer(device=device)
        self.loss = ref.PolicyLoss()
        # materialise the LazyLinear layers, then load the same initial weights our arm uses
        self.step()
        flat = WL.init_params(widths, torch.Generator().manual_seed(0), "cpu")
        sd, o = self.model.state_dict(), 0
        for i in range(len(widths) - 1):
            n = widths[i + 1] * widths[i]
            sd[f"net.master.{2 * i}.weight"].copy_(flat[o:o + n].view(widths[i + 1], widths[i]))
            o += n
            sd[f"net.master.{2 * i}.bias"].copy_(flat[o:o + widths[i + 1]])
            o += widths[i + 1]

    def step(self):
        """trainer.py:160-173 for one batch: zero grads, simulate, mean loss, backward. Returns the total cost."""
        for prm in self.model.parameters():
            prm.grad = None
        data = {k: v.clone() for k, v in self.data.items()}  # reset() keeps references into the batch
        total, report = self.trainer.simulate_batch(self.loss, self.sim, self.model, self.T, self.problem_params, data,
                                                    self.obs_params, 30, False)
        (total / (self.B * self.T * self.S)).backward()
        return total

    def time(self, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        if self.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            tot = self.step()
        if self.device != "cpu":
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        return self.B * self.T / dt, dt, float(tot.detach())
