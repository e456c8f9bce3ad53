"""Trainer with the reference's interface (trainer.py:5-324); `simulate_batch` is the drop-in boundary.

`simulate_batch(loss_function, simulator, model, periods, problem_params, data_batch, observation_params,
ignore_periods=0, discrete_allocation=False) -> (batch_reward, reward_to_report)` returns two 0-d tensors that
carry an autograd graph, exactly like the reference - but for fusable policies (+ PolicyLoss) the graph is ONE
node: the whole T-period rollout ran in the fused forward kernel and `mean_loss.backward()` triggers the
reverse-time adjoint kernel. Everything else in `do_one_epoch` / `train` / `test` (optimizer step, gradient
clipping, best-model tracking, early stopping, checkpoints) is the reference's own control flow.
"""
import copy
import datetime
import os
import weakref

import numpy as np
import torch

from . import device_dataset as DD
from . import engine as EN
from . import parallel as PL
from .loss_functions import PolicyLoss

_FUSED_MODULE_ORDER = ("master", "context", "store", "warehouse")


class Trainer:
    def __init__(self, device="cpu"):
        self.all_train_losses = []
        self.all_dev_losses = []
        self.all_test_losses = []
        self.device = device
        self.time_stamp = self.get_time_stamp()
        self.best_performance_data = {"train_loss": np.inf, "dev_loss": np.inf, "last_epoch_saved": -1000,
                                      "model_params_to_save": None}
        self.best_epoch = 0
        self.use_fused = os.environ.get("HDPO_DISABLE_FUSED", "0") != "1"
        # matmul precision request for policies with a tensor-core rollout (the wide MLPs): "tf32x3" is the
        # parity-grade tcgen05 mode, "fp32" the SIMT FFMA mode, "tf32" single-pass (throughput, not parity-grade)
        self.precision = os.environ.get("HDPO_PRECISION", "tf32x3")
        self._engines = {}
        # keep datasets resident in HBM and assemble batches with a gather kernel instead of per-sample collate
        self.use_device_dataset = os.environ.get("HDPO_DEVICE_DATASET", "1") != "0"
        self._device_loaders = {}
        self.last_path = None  # "fused" | "generic": which path the last simulate_batch took (for tests/logging)
        self._replicas_synced = False  # multi-rank: parameters broadcast from rank 0 once they are materialised
        # fused training step: flat parameter / gradient / Adam-moment vectors, adjoint written straight
        # into the flat gradient, ONE fused Adam kernel, no host synchronisation inside the batch loop
        self.use_fused_step = os.environ.get("HDPO_FUSED_STEP", "1") != "0"
        self._flat_ctx = {}

    def reset(self):
        self.all_train_losses, self.all_dev_losses, self.all_test_losses = [], [], []

    # ------------------------------------------------------------------ epoch loops (reference control flow)
    def train(self, epochs, loss_function, simulator, model, data_loaders, optimizer, problem_params,
              observation_params, params_by_dataset, trainer_params):
        for epoch in range(epochs):
            tr = params_by_dataset["train"]
            _, train_report = self.do_one_epoch(optimizer, data_loaders["train"], loss_function, simulator, model,
                                                tr["periods"], problem_params, observation_params, train=True,
                                                ignore_periods=tr["ignore_periods"])
            self.all_train_losses.append(train_report)
            if epoch % trainer_params["do_dev_every_n_epochs"] == 0:
                dv = params_by_dataset["dev"]
                _, dev_report = self.do_one_epoch(optimizer, data_loaders["dev"], loss_function, simulator, model,
                                                  dv["periods"], problem_params, observation_params, train=False,
                                                  ignore_periods=dv["ignore_periods"])
                self.all_dev_losses.append(dev_report)
                self.update_best_params_and_save(epoch, train_report, dev_report, trainer_params, model, optimizer)
                patience = trainer_params.get("early_stopping_patience_epochs", None)
                if patience is not None and (epoch - self.best_epoch) >= patience:
                    print(f"\nEarly stopping triggered at epoch {epoch + 1}")
                    print(f"No improvement for {epoch - self.best_epoch} epochs")
                    print(f"Best model was at epoch {self.best_epoch + 1} with dev loss: "
                          f"{self.best_performance_data['dev_loss']}")
                    break
            else:
                dev_report = 0
                self.all_dev_losses.append(self.all_dev_losses[-1])
            if epoch % trainer_params["print_results_every_n_epochs"] == 0:
                print(f"epoch: {epoch + 1}")
                print(f"Average per-period train loss: {train_report}")
                print(f"Average per-period dev loss: {dev_report}")
                print(f"Best per-period dev loss: {self.best_performance_data['dev_loss']}")

    def test(self, loss_function, simulator, model, data_loaders, optimizer, problem_params, observation_params,
             params_by_dataset, discrete_allocation=False):
        if model.trainable and self.best_performance_data["model_params_to_save"] is not None:
            model.load_state_dict(self.best_performance_data["model_params_to_save"])
        ts = params_by_dataset["test"]
        return self.do_one_epoch(optimizer, data_loaders["test"], loss_function, simulator, model, ts["periods"],
                                 problem_params, observation_params, train=False,
                                 ignore_periods=ts["ignore_periods"], discrete_allocation=discrete_allocation)

    def do_one_epoch(self, optimizer, data_loader, loss_function, simulator, model, periods, problem_params,
                     observation_params, train=True, ignore_periods=0, discrete_allocation=False):
        # epoch sums live on the device: the reference's two .item() reads per batch (trainer.py:166-167) are host
        # synchronisations that keep the CPU from running ahead of the GPU; they happen ONCE per epoch here
        epoch_loss = None
        epoch_loss_to_report = None
        total_samples = len(data_loader.dataset)
        n_stores = problem_params["n_stores"]
        rank, world = PL.world_info()  # (0, 1) unless launched under torchrun with an initialised process group
        data_loader = self._maybe_device_loader(data_loader)
        with torch.no_grad() if not train else torch.enable_grad():
            for data_batch in data_loader:
                n_global = len(data_batch["demands"])
                # scenario-parallel: every rank sees the same batch order (same sampler seed) and keeps its shard
                data_batch = self.move_batch_to_device(PL.shard_batch(data_batch, rank, world))
                if world > 1 and not self._replicas_synced:
                    # replicas start from rank 0's weights (materialise the lazy layers first, as the first batch would)
                    self._materialize(model, simulator, periods, problem_params, data_batch, observation_params)
                    PL.broadcast_parameters_(model)
                    self._replicas_synced = True
                fast = None
                if train and model.trainable and not discrete_allocation:
                    fast = self._fused_step_ctx(optimizer, loss_function, simulator, model, periods, problem_params,
                                                data_batch, observation_params)
                if fast is not None:
                    total_reward, reward_to_report = self._train_batch_fused(
                        fast, optimizer, simulator, model, periods, problem_params, data_batch, observation_params,
                        ignore_periods, n_global, world)
                else:
                    if train:
                        optimizer.zero_grad()
                    total_reward, reward_to_report = self.simulate_batch(
                        loss_function, simulator, model, periods, problem_params, data_batch, observation_params,
                        ignore_periods, discrete_allocation)
                    mean_loss = total_reward / (n_global * periods * n_stores)
                    do_step = train and model.trainable
                    if do_step:
                        mean_loss.backward()
                    if world > 1:  # ONE collective per batch: flat gradient + the two loss scalars
                        total_reward, reward_to_report = total_reward.detach().clone(), reward_to_report.detach().clone()
                        PL.allreduce_sum_(([p.grad for p in model.parameters() if p.grad is not None] if do_step else [])
                                          + [total_reward, reward_to_report])
                    if do_step:
                        clip = getattr(model, "gradient_clipping_norm_value", None)
                        if clip is not None:
                            torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
                        optimizer.step()
                tr_, rr_ = torch.as_tensor(total_reward).detach(), torch.as_tensor(reward_to_report).detach()
                epoch_loss = tr_.double() if epoch_loss is None else epoch_loss + tr_.double()
                epoch_loss_to_report = rr_.double() if epoch_loss_to_report is None else epoch_loss_to_report + rr_.double()
        epoch_loss = 0.0 if epoch_loss is None else float(epoch_loss)
        epoch_loss_to_report = 0.0 if epoch_loss_to_report is None else float(epoch_loss_to_report)
        return (epoch_loss / (total_samples * periods * n_stores),
                epoch_loss_to_report / (total_samples * (periods - ignore_periods) * n_stores))

    # ------------------------------------------------------------------ fused training step
    def _fused_step_ctx(self, optimizer, loss_function, simulator, model, periods, problem_params, data_batch,
                        observation_params):
        """Flat-vector context when this (model, optimizer) pair can take the fused step: a fusable policy on the stock
        simulator with PolicyLoss, and a plain torch.optim.Adam (one parameter group, no amsgrad / maximize /
        capturable) over exactly the policy's parameters. Otherwise None (reference control flow)."""
        if not self.use_fused_step:
            return None
        pspec = self._fusable_spec(loss_function, simulator, model, periods, problem_params, data_batch,
                                   observation_params)
        if pspec is None or type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
            return None
        grp = optimizer.param_groups[0]
        if grp.get("amsgrad") or grp.get("maximize") or grp.get("capturable") or grp.get("differentiable"):
            return None
        key = (id(model), id(optimizer))
        ctx = self._flat_ctx.get(key)
        if ctx is not None and ctx["model_ref"]() is model and ctx["opt_ref"]() is optimizer:
            ctx["pspec"] = pspec
            return ctx
        names = [m for m in _FUSED_MODULE_ORDER if m in model.net]
        params = [p for m in names for p in model.net[m].parameters()]
        if {id(p) for p in params} != {id(p) for p in grp["params"]} or len(params) != len(list(model.parameters())):
            return None
        if any(p.dtype != torch.float32 or not p.is_cuda for p in params):
            return None
        n = sum(p.numel() for p in params)
        dev = params[0].device
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        grad = torch.zeros(n, dtype=torch.float32, device=dev)
        m1 = torch.zeros(n, dtype=torch.float32, device=dev)
        m2 = torch.zeros(n, dtype=torch.float32, device=dev)
        step, o = 0, 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat[o:o + k].copy_(p.detach().reshape(-1))
                st = optimizer.state.get(p, {})
                if "exp_avg" in st:  # resumed from a checkpoint (Trainer.load_model): keep the moments
                    m1[o:o + k].copy_(st["exp_avg"].reshape(-1))
                    m2[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
                    step = int(st["step"])
                # parameters, gradients and moments become VIEWS of the flat vectors: state_dict(), checkpoints,
                # clip_grad_norm_ and any torch code keep working on the same memory
                p.data = flat[o:o + k].view_as(p)
                p.grad = grad[o:o + k].view_as(p)
                optimizer.state[p] = {"step": torch.tensor(float(step)), "exp_avg": m1[o:o + k].view_as(p),
                                      "exp_avg_sq": m2[o:o + k].view_as(p)}
                o += k
        ctx = {"flat": flat, "grad": grad, "m1": m1, "m2": m2, "step": step, "params": params, "pspec": pspec,
               "model_ref": weakref.ref(model), "opt_ref": weakref.ref(optimizer)}
        if len(self._flat_ctx) > 4:
            self._flat_ctx.clear()
        self._flat_ctx[key] = ctx
        return ctx

    def _train_batch_fused(self, ctx, optimizer, simulator, model, periods, problem_params, data_batch, observation_params,
                           ignore_periods, n_global, world):
        """forward rollout -> adjoint with dLoss/dtotal = 1 / (B_global T S) (trainer.py:169) into the flat gradient ->
        [all-reduce] -> [clip] -> fused Adam; returns the two loss scalars as device tensors (nothing is read back)."""
        eng = self._engine_for(ctx["pspec"], model, periods, problem_params, data_batch, observation_params,
                               ignore_periods, False, True)
        if simulator.observation is None or getattr(simulator, "batch_size", None) != len(data_batch["demands"]):
            simulator.reset(periods, problem_params, data_batch, observation_params)  # attributes a caller may inspect
        o = 0
        for p in ctx["params"]:  # optimizer.zero_grad(set_to_none=True) elsewhere may have dropped the gradient views
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != ctx["grad"].data_ptr() + 4 * o:
                p.grad = ctx["grad"][o:o + k].view_as(p)
            o += k
        totals = eng.forward(ctx["flat"], data_batch)
        n_stores = problem_params["n_stores"]
        eng.backward(1.0 / (n_global * periods * n_stores), 0.0, out=ctx["grad"])
        totals32 = totals.to(torch.float32)
        if world > 1:
            totals32 = totals32.clone()
            PL.allreduce_sum_([ctx["grad"], totals32])
        clip = getattr(model, "gradient_clipping_norm_value", None)
        if clip is not None:  # torch.nn.utils.clip_grad_norm_ on the flat vector, without its host read
            coef = torch.clamp(float(clip) / (ctx["grad"].norm() + 1e-6), max=1.0)
            ctx["grad"].mul_(coef)
        grp = optimizer.param_groups[0]
        # the step count lives in the optimizer's own (CPU) state tensors, so torch code that touches the optimizer in
        # between (a generic-path batch, load_state_dict) stays consistent with this path
        ctx["step"] = int(optimizer.state[ctx["params"][0]]["step"]) + 1
        rc = eng.lib.hdpo_adam_step(ctx["flat"].data_ptr(), ctx["grad"].data_ptr(), ctx["m1"].data_ptr(),
                                    ctx["m2"].data_ptr(), ctx["flat"].numel(), float(grp["lr"]), float(grp["betas"][0]),
                                    float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]), ctx["step"],
                                    EN.current_stream_ptr(ctx["flat"].device))
        EN.K.check(eng.lib, rc, "hdpo_adam_step")
        for p in ctx["params"]:
            optimizer.state[p]["step"].fill_(float(ctx["step"]))  # CPU scalar tensors: no device round trip
        self.last_path = "fused"
        return totals32[0], totals32[1]

    def _maybe_device_loader(self, data_loader):
        if not self.use_device_dataset or not str(self.device).startswith("cuda") or not DD.eligible(data_loader):
            return data_loader
        key = id(data_loader)
        if key not in self._device_loaders:
            self._device_loaders[key] = DD.DeviceBatches(data_loader, self.device)
        return self._device_loaders[key]

    # ------------------------------------------------------------------ the hot path
    def simulate_batch(self, loss_function, simulator, model, periods, problem_params, data_batch, observation_params,
                       ignore_periods=0, discrete_allocation=False):
        pspec = self._fusable_spec(loss_function, simulator, model, periods, problem_params, data_batch,
                                   observation_params)
        if pspec is not None:
            try:
                out = self._simulate_batch_fused(pspec, simulator, model, periods, problem_params, data_batch,
                                                 observation_params, ignore_periods, discrete_allocation)
                self.last_path = "fused"
                return out
            except EN.K.HdpoError as e:
                if "no fused rollout" not in str(e):
                    raise
        self.last_path = "generic"
        return self._simulate_batch_generic(loss_function, simulator, model, periods, problem_params, data_batch,
                                            observation_params, ignore_periods, discrete_allocation)

    def _simulate_batch_generic(self, loss_function, simulator, model, periods, problem_params, data_batch,
                                observation_params, ignore_periods, discrete_allocation):
        """Any torch policy: policy(obs) in torch, one K3 kernel per period (trainer.py:181-216 semantics)."""
        batch_reward = 0
        reward_to_report = 0
        observation, _ = simulator.reset(periods, problem_params, data_batch, observation_params)
        for t in range(periods):
            obs_and_internal = dict(observation)
            obs_and_internal["internal_data"] = simulator._internal_data  # for non-admissible benchmark policies
            action = model(obs_and_internal)
            if discrete_allocation:
                action = {key: val.round() for key, val in action.items()}
            observation, reward, terminated, _, _ = simulator.step(action)
            total_reward = loss_function(None, action, reward)
            batch_reward += total_reward
            if t >= ignore_periods:
                reward_to_report += total_reward
            if terminated:
                break
        return batch_reward, reward_to_report

    def _fusable_spec(self, loss_function, simulator, model, periods, problem_params, data_batch, observation_params):
        if not self.use_fused or type(loss_function) is not PolicyLoss or not hasattr(model, "fusable_spec"):
            return None
        if not data_batch["initial_inventories"].is_cuda:
            return None
        # the kernels implement the STOCK dynamics and the registered policies' own forward: a subclass that overrides
        # either (custom simulator, custom forward) must not be silently bypassed - it runs on the generic path
        from .environment import Simulator as _StockSimulator
        from .neural_networks import NeuralNetworkCreator as _Creator
        if type(simulator) is not _StockSimulator:
            return None
        registered = None
        try:
            registered = _Creator().get_architecture(model.nn_args["name"])
        except Exception:  # noqa: BLE001
            registered = None
        if registered is None or type(model).forward is not registered.forward:
            return None
        pspec = model.fusable_spec()
        if pspec is None and self._materialize(model, simulator, periods, problem_params, data_batch, observation_params):
            pspec = model.fusable_spec()
        if pspec is not None and pspec.arch == "symmetry_aware" and not ("mean" in data_batch and "std" in data_batch):
            return None  # the generic path raises the same KeyError the reference's forward would
        return pspec

    def _materialize(self, model, simulator, periods, problem_params, data_batch, observation_params):
        """Materialise LazyLinear layers with one throw-away single-scenario forward (initialisation only: the draws
        from the global RNG are the ones the first real batch would make). Returns True when something was lazy."""
        if not any(isinstance(m, torch.nn.modules.lazy.LazyModuleMixin) and m.has_uninitialized_params()
                   for m in model.modules()):
            return False
        one = {k: v[:1] for k, v in data_batch.items()}
        with torch.no_grad():
            obs, _ = simulator.reset(periods, problem_params, one, observation_params)
            obs = dict(obs)
            obs["internal_data"] = simulator._internal_data
            model(obs)
        return True

    def _flat_params(self, model, pspec):
        names = [m for m in _FUSED_MODULE_ORDER if m in model.net]
        params = [p for m in names for p in model.net[m].parameters()]
        if len(params) == 1:
            return params[0].reshape(-1)
        return torch.cat([p.reshape(-1) for p in params])

    def _simulate_batch_fused(self, pspec, simulator, model, periods, problem_params, data_batch, observation_params,
                              ignore_periods, discrete_allocation):
        need_grad = torch.is_grad_enabled() and model.trainable and not discrete_allocation
        eng = self._engine_for(pspec, model, periods, problem_params, data_batch, observation_params, ignore_periods,
                               discrete_allocation, need_grad)
        # keep the simulator's visible state coherent with what a per-period run would leave behind
        simulator.reset(periods, problem_params, data_batch, observation_params)
        flat = self._flat_params(model, pspec)
        if need_grad:
            total, report = EN.rollout(eng, data_batch, flat)
        else:
            totals = eng.forward(flat.detach(), data_batch).to(torch.float32)
            total, report = totals[0], totals[1]
        simulator.observation["current_period"] += periods
        return total, report

    def _engine_for(self, pspec, model, periods, problem_params, data_batch, observation_params, ignore_periods,
                    discrete_allocation, need_grad):
        """The cached FusedRollout (descriptor + workspace) of this (model, problem, batch shape)."""
        shift = observation_params["demand"]["period_shift"]
        B = data_batch["initial_inventories"].shape[0]
        # everything the cached descriptor bakes in: batch / state shapes, problem flags, adjacency, policy widths and
        # activations, precision (the same model object may be reused with other problem_params)
        shapes = tuple((k, tuple(v.shape[1:])) for k, v in sorted(data_batch.items()) if k.startswith("initial_"))
        adj = pspec.adjacency
        adj_key = None if adj is None else tuple(map(tuple, adj.tolist() if hasattr(adj, "tolist") else adj))
        nets = tuple((tuple(n[0]), n[1], n[2]) if n is not None else None
                     for n in (pspec.master, pspec.store_net, pspec.warehouse_net))
        key = (id(model), B, periods, data_batch["demands"].shape[2], ignore_periods, bool(discrete_allocation),
               need_grad, shift, pspec.warehouse_upper_bound, shapes, adj_key, nets, pspec.transshipment, pspec.prop_eps,
               self.precision, int(problem_params["n_stores"]), int(problem_params["n_warehouses"]),
               int(problem_params["n_extra_echelons"]), bool(problem_params["lost_demand"]),
               bool(problem_params["maximize_profit"]), "warehouse_edge_costs" in data_batch)
        eng = self._engines.get(key)
        if eng is not None and eng.model_ref() is not model:  # id() of a collected model reused by a new object
            eng = None
        if eng is None:
            if len(self._engines) > 8:
                self._engines.clear()
            eng = EN.FusedRollout(pspec, problem_params, data_batch, periods, ignore_periods=ignore_periods,
                                  period_shift=shift, discrete_allocation=discrete_allocation,
                                  save_for_backward=need_grad, precision=self.precision)
            eng.model_ref = weakref.ref(model)
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------ checkpoints / bookkeeping
    def save_model(self, epoch, model, optimizer, trainer_params):
        path = self.create_many_folders_if_not_exist_and_return_path(trainer_params["base_dir"],
                                                                     trainer_params["save_model_folders"])
        # key names (incl. the reference's mislabelled ones, trainer.py:227-229) are kept for file compatibility
        torch.save({
            "epoch": epoch,
            "model_state_dict": self.best_performance_data["model_params_to_save"],
            "optimizer_state_dict": optimizer.state_dict(),
            "best_train_loss": self.best_performance_data["dev_loss"],
            "best_dev_loss": self.all_train_losses,
            "all_train_losses": self.all_train_losses,
            "all_dev_losses": self.all_dev_losses,
            "all_test_losses": self.all_test_losses,
            "warehouse_upper_bound": model.warehouse_upper_bound,
        }, f"{path}/{trainer_params['save_model_filename']}.pt")

    def create_folder_if_not_exists(self, folder):
        if not os.path.isdir(folder):
            os.mkdir(folder)

    def create_many_folders_if_not_exist_and_return_path(self, base_dir, intermediate_folder_strings):
        path = base_dir
        for part in intermediate_folder_strings:
            path += f"/{part}"
            self.create_folder_if_not_exists(path)
        return path

    def update_best_params_and_save(self, epoch, train_loss, dev_loss, trainer_params, model, optimizer):
        current = {"train_loss": train_loss, "dev_loss": dev_loss}
        metric = trainer_params["choose_best_model_on"]
        if current[metric] < self.best_performance_data[metric]:
            self.best_performance_data["train_loss"] = train_loss
            self.best_performance_data["dev_loss"] = dev_loss
            if model.trainable:
                self.best_performance_data["model_params_to_save"] = copy.deepcopy(model.state_dict())
            self.best_performance_data["update"] = True
            self.best_epoch = epoch
        if trainer_params["save_model"] and model.trainable:
            due = self.best_performance_data["last_epoch_saved"] + trainer_params["epochs_between_save"] <= epoch
            if due and self.best_performance_data["update"]:
                self.best_performance_data["last_epoch_saved"] = epoch
                self.best_performance_data["update"] = False
                self.save_model(epoch, model, optimizer, trainer_params)

    def plot_losses(self, ymin=None, ymax=None):
        from .shared_imports import plt
        if plt is None:
            raise RuntimeError("matplotlib is not installed")
        plt.plot(self.all_train_losses, label="Train loss")
        plt.plot(self.all_dev_losses, label="Dev loss")
        plt.legend()
        if ymin is not None and ymax is not None:
            plt.ylim(ymin, ymax)
        plt.xlabel("Epoch")
        plt.ylabel("Loss")
        plt.show()

    def move_batch_to_device(self, data_batch):
        return {k: v.to(self.device, non_blocking=True) for k, v in data_batch.items()}

    def load_model(self, model, optimizer, model_path):
        checkpoint = torch.load(model_path, map_location=self.device, weights_only=False)
        model.load_state_dict(checkpoint["model_state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        self.all_train_losses = checkpoint["all_train_losses"]
        self.all_dev_losses = checkpoint["all_dev_losses"]
        self.all_test_losses = checkpoint["all_test_losses"]
        model.warehouse_upper_bound = checkpoint["warehouse_upper_bound"]
        return model, optimizer

    def get_time_stamp(self):
        return int(datetime.datetime.now().timestamp())

    def get_year_month_day(self):
        now = datetime.datetime.now()
        return f"{now.year}_{now.month:02d}_{now.day:02d}"
