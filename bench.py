#!/usr/bin/env python
"""bench.py - train scenario-periods/s (forward + adjoint rollout) of the HDPO hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one fused forward rollout + reverse-time adjoint over one batch of synthetic demand that is already
resident in HBM (plus, for N > 1, the policy-gradient all-reduce over NCCL). Scenario batches are sharded across
ranks with a fixed per-GPU batch (weak scaling). Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
`--impl reference` times the UNMODIFIED reference (its own Trainer.simulate_batch + backward, imported from
baseline/_ref, see baseline/ref_harness.py) on the host cores, all threads, on a bounded sample of the same workload
(its PyTorch-eager port, oracle/torch_port.py, only where the reference snapshot has no such policy).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json names no config for its metric; its north-star target ("fused forward+adjoint rollouts on
# one_warehouse_lost_demand at 1 GPU") does, so that is the default workload: configs[3] with the shipped
# vanilla_warehouse policy (153 -> 512^3 -> 51), 8192 scenarios x 50 periods x 50 stores (fits one GPU: ~12 GB).
# configs[0..2] and [4] are selected with --workload (numbers in profiles/README.md).
DEFAULT_WORKLOAD = "one_warehouse_lost_demand"
METRIC = "train scenario-periods/sec (fwd+bwd)"
UNIT = "scenario-periods/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_traffic(workload):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed ncu --set full
    captures (profiles/r2_traffic.json names the capture each number comes from); None when there is none."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


def time_dominant_gemm(lib, torch, dev, stream, M, N, K, n_pass, reps=64):
    """The dominant kernel of the wide rollout timed ALONE: gemm_tc_kernel (tcgen05 / TMEM / TMA, CTA-pair tiles) in
    its forward hidden-layer form (bias + ELU epilogue, (hi, lo) outputs) at the shape one scenario chunk launches.
    CUDA events on the launching stream; operands rotate through 16 buffer sets (> L2) so that no launch finds its
    inputs in L2 from the launch before."""
    nbuf = 16
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.randn(M, K, generator=g, device=dev)
    Bm = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    bias = torch.zeros(N, device=dev)
    sets = []
    for _ in range(nbuf):
        scratch = torch.empty(2 * (M * K + N * K), device=dev)
        Cm, Clo = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
        rc = lib.hdpo_debug_gemm_tc(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, n_pass, scratch.data_ptr(), stream)
        assert rc == 0, lib.hdpo_last_error()
        sets.append((scratch, Cm, Clo))
    dbg = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)

    def run(i):
        scratch, Cm, Clo = sets[i % nbuf]
        rc = lib.hdpo_debug_gemm_tc_timeline(A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), M, N, K, n_pass,
                                             scratch.data_ptr(), dbg.data_ptr(), stream, 0, Clo.data_ptr(),
                                             bias.data_ptr())
        assert rc == 0, lib.hdpo_last_error()

    for i in range(nbuf):
        run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3  # seconds per launch


def cpu_port_setup(workload, B, T, seed=0):
    """The same synthetic workload on the host, as inputs of the PyTorch-eager port (oracle/torch_port.py)."""
    import torch
    from neural_inventory_control_b200 import workloads as WL
    pspec, pp, data, widths = WL.WORKLOADS[workload]("cpu", B=B, T=T)
    g = torch.Generator().manual_seed(seed)
    flat = WL.init_params(widths, g, "cpu")
    nets, o = {}, 0
    for name, ws in WL.net_list(widths):
        layers = []
        for i in range(len(ws) - 1):
            n = ws[i + 1] * ws[i]
            w = flat[o:o + n].view(ws[i + 1], ws[i]).clone().requires_grad_(True)
            o += n
            b = flat[o:o + ws[i + 1]].clone().requires_grad_(True)
            o += ws[i + 1]
            layers.append((w, b))
        nets[name] = layers
    first = nets["master"] if "master" in nets else nets["context"]
    pol = {"arch": pspec.arch, "layers": [wb for layers in nets.values() for wb in layers],
           "hidden_act": pspec.master[1], "out_act": pspec.master[2],
           "wub": torch.tensor([pspec.warehouse_upper_bound]),
           "adjacency": None if pspec.adjacency is None else torch.tensor(pspec.adjacency),
           "transshipment": pspec.transshipment}
    if pspec.arch == "symmetry_aware":
        pol["nets"] = {"context": (first, pspec.master[1], pspec.master[2]),
                       "store": (nets["store"], pspec.store_net[1], pspec.store_net[2]),
                       "warehouse": (nets["warehouse"], pspec.warehouse_net[1], pspec.warehouse_net[2])}
        pol["prop_eps"] = pspec.prop_eps
    pb = dict(pp, period_shift=0)
    return pol, pb, data


def time_cpu_port(workload, B, T, steps, warmup):
    import torch
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pol, pb, data = cpu_port_setup(workload, B, T)
    for _ in range(warmup):
        TP.train_step(pol, pb, data, T)
    t0 = time.perf_counter()
    for _ in range(steps):
        TP.train_step(pol, pb, data, T)
    dt = (time.perf_counter() - t0) / steps
    return B * T / dt, dt, cores


def cpu_sample_size(workload):
    """Scenarios per step of the CPU arm: the reference's own training batch where the host can afford it
    (vanilla_warehouse.yml / one_store YAMLs: 1024 / 8192 per batch), a bounded sample of the 2^20 workloads."""
    return {"one_store_backlogged_lead20": 16384, "one_store_lost": 8192, "serial_system": 8192,
            "one_warehouse_lost_demand": 1024, "many_warehouses_lost_demand": 1024,
            "one_warehouse_lost_demand_symmetry_aware": 512}.get(workload, 1024)


def time_reference(workload, B, T, steps, warmup, device="cpu"):
    """(value, seconds per step, cores, kind, note): the unmodified reference when it can be imported and has the
    policy, else its PyTorch-eager port."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        from baseline import ref_harness as RH
        run = RH.ReferenceRun(workload, B, T, device=device)
        value, dt, _ = run.time(steps, warmup)
        return value, dt, cores, "reference", f"unmodified reference from {os.path.relpath(run.where, ROOT)}"
    except (ImportError, NotImplementedError) as e:
        if device != "cpu":
            raise
        value, dt, cores = time_cpu_port(workload, B, T, steps, warmup)
        return value, dt, cores, "port", f"PyTorch-eager port ({e})"


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Bc, T = args.batch or cpu_sample_size(args.workload), args.periods
    value, dt, cores, kind, note = time_reference(args.workload, Bc, T, args.steps, args.warmup)
    sample = (f"{Bc} scenarios x {T} periods per step (bounded sample of the workload: the reference's own training batch "
              f"size where the host affords it), {note}, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "periods": T, "scenarios_per_step": Bc, "device": "host cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--batch", type=int, default=None, help="scenarios per GPU (default: the workload's)")
    ap.add_argument("--periods", type=int, default=50)
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32x3", "tf32"],
                    help="policy-MLP matmul mode (default: tf32x3 = parity-grade tcgen05 where available, else fp32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true",
                    help="skip the short device-timed runs of the other BASELINE configs reported under 'other_workloads'")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from neural_inventory_control_b200 import _capi as K
    from neural_inventory_control_b200 import engine as EN
    from neural_inventory_control_b200 import workloads as WL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the HDPO engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            # keep NCCL's version banner (printed at VERSION and WARN level) off stdout: rank 0 prints ONE JSON line
            os.environ.pop("NCCL_DEBUG", None)
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
        dist.init_process_group("nccl", device_id=dev)

    kw = {"T": args.periods}
    if args.batch:
        kw["B"] = args.batch
    pspec, pp, data, widths = WL.WORKLOADS[args.workload](dev, seed=57 + rank, **kw)
    B, S, T = data["demands"].shape[0], pp["n_stores"], args.periods
    gen = torch.Generator(device=dev).manual_seed(0)  # identical weights on every rank
    flat = WL.init_params(widths, gen, dev)
    small = pspec.arch in ("vanilla_one_store", "vanilla_serial")
    # tf32x3 everywhere: tcgen05 GEMMs in the wide path, warp-level mma.sync in the adjoint of the small nets and of the
    # SymmetryAware heads - all with the fp32-grade 3xTF32 split (parity-tested at the same 1e-5 bar as fp32)
    precision = args.precision or "tf32x3"
    eng = EN.FusedRollout(pspec, pp, data, T, ignore_periods=30, precision=precision)
    grad = torch.zeros_like(flat)
    lib = eng.lib
    g_total = 1.0 / (world * B * T * S)

    def step():
        eng.forward(flat, data)
        eng.backward(g_total, 0.0, out=grad)
        if world > 1:
            dist.all_reduce(grad)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.hdpo_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.hdpo_kernel_launch_count() - l0
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B * T / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (the adjoint), CUDA events on the launching stream
    def time_region(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    n_k = max(3, min(args.steps, 10))
    fwd_ms = time_region(lambda: eng.forward(flat, data), n_k)
    bwd_ms = time_region(lambda: eng.backward(g_total, 0.0, out=grad), n_k)
    peaks = load_peaks()
    macs = WL.forward_macs(widths, S)
    detached = pspec.arch == "vanilla_serial"
    flops_step = WL.flops_per_scenario_period(widths, first_layer_dgrad=not detached, n_stores=S)
    sym = pspec.arch == "symmetry_aware"
    flops_bwd = flops_step - 2 * macs  # dgrad + wgrad; the adjoint kernel's recompute is not counted
    tf32_peak = peaks["bf16_tflops"] / 2.0
    achieved = flops_bwd * B * T / (bwd_ms * 1e-3) / 1e12
    traffic = load_traffic(args.workload)
    group = {
        "kernel": (("small_bwd_kernel (reverse-time adjoint; HxH layers on mma.sync 3xTF32, first / output layer and simulator in "
                    "fp32 FFMA)" if precision != "fp32" else "small_bwd_kernel (reverse-time adjoint, SIMT fp32 parity mode)")
                   if small else
                   f"adjoint sweep ({precision}): sym_head_bwd_kernel (store / warehouse nets recomputed + adjoint, SIMT "
                   "fp32) + gemm_tc_kernel dgrad / weight-gradient tiles of the context trunk (tcgen05/TMEM/TMA)" if sym
                   and precision != "fp32" else
                   f"adjoint sweep ({precision}): gemm_tc_kernel dgrad + weight-gradient tiles (tcgen05/TMEM/TMA) + "
                   "warehouse_head_bwd + bias column sums" if precision != "fp32" else
                   "adjoint sweep: sgemm_kernel dgrad+wgrad tiles (SIMT fp32 parity mode) + warehouse_head_bwd"),
        "achieved": achieved, "frac": achieved / tf32_peak, "kernel_ms": bwd_ms, "fwd_kernel_ms": fwd_ms,
        "flops_per_launch": flops_bwd * B * T,
    }
    roofline = {
        "bound": "tensor", "kernel": group["kernel"], "precision": precision,
        "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
        "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
        "traffic_source": traffic["source"] if traffic else None,
        "peak_source": f"{peaks['source']}: tf32 taken as bf16_tflops/2 (burst, kernel timed alone)",
        "kernel_ms": bwd_ms, "fwd_kernel_ms": fwd_ms, "flops_per_launch": flops_bwd * B * T,
        "step_frac": flops_step * B * T / (ms_per_step * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] / 2.0),
        "hbm_frac": 8.0 * S * B * T / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
    }
    if pspec.arch == "vanilla_warehouse" and precision != "fp32":
        # dominant single kernel (56 % of the step in the ncu launch list): the hidden-layer tile GEMM, timed alone at
        # the shape one scenario chunk launches; the whole adjoint group stays reported beside it
        n_chunks = int(os.environ.get("HDPO_WIDE_CHUNKS", "0")) or max(1, min(4, B // 2048))
        per_chunk = -(-B // n_chunks)
        Mc = -(-per_chunk // 128) * 128  # rows of one chunk, padded to the 128-row tile
        hidden = widths[1]
        n_pass = 3 if precision == "tf32x3" else 1
        sec = time_dominant_gemm(lib, torch, dev, EN.current_stream_ptr(dev), Mc, hidden, hidden, n_pass)
        fl = 2.0 * Mc * hidden * hidden
        # Headline fraction = the WHOLE STEP (every launch of the timed region: algorithmic FLOPs of forward + dgrad +
        # wgrad over the device-timed step, against the sustained tf32 peak - the kernels run inside a long step); the
        # dominant kernel timed alone (one chunk launch = 64 of 148 SMs by construction, burst peak) explains it.
        sustained = peaks["bf16_tflops_sustained"] / 2.0
        step_tf = flops_step * B * T / (ms_per_step * 1e-3) / 1e12
        roofline.update({
            "kernel": f"gemm_tc_kernel<128, FWD_HIDDEN, CTA pair> (tcgen05 kind::tf32 {precision}, TMEM accumulators, "
                      f"TMA operands): [{Mc} x {hidden}] x [{hidden} x {hidden}] + bias + ELU + (hi, lo) split",
            "achieved": step_tf, "peak": sustained, "frac": step_tf / sustained,
            "peak_source": f"{peaks['source']}: tf32 taken as bf16_tflops_sustained / 2 (kernels timed inside a long step)",
            "frac_basis": "whole step: algorithmic FLOPs of all launches of the timed region / device-timed step, "
                          "against the measured sustained tf32 peak (bf16 / 2)",
            "kernel_alone": {"achieved": fl / sec / 1e12, "peak": tf32_peak, "frac": fl / sec / 1e12 / tf32_peak,
                             "flops_per_launch": fl, "launch_us": sec * 1e6,
                             "peak_source": "measured burst bf16 / 2 (kernel timed alone, operands rotating through "
                                            "16 buffer sets > L2)"},
            "flops_per_launch": fl, "launch_us": sec * 1e6,
            "launches_per_step": 2 * (len(widths) - 3) * n_chunks * T,
            "note": "algorithmic FLOPs: the 3 tensor passes of the fp32-grade split count once (ceiling 1/3)"
                    if n_pass == 3 else "single tensor pass",
            "adjoint_group": group,
        })

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the batch + D2H of loss and gradient.
    # Headline e2e: the synthetic demand is generated ON THE DEVICE inside the call (K4 Philox sampler wired into the
    # path: demands = NULL + seed / offset in the descriptor), so what crosses the bus per step is parameters, initial
    # inventories, cost coefficients and S means / standard deviations in, totals + gradient out. The variant that
    # uploads a host-resident demand tensor every step stays reported beside it (e2e_host_demand).
    e2e = None
    e2e_host_demand = None
    if not args.no_e2e:
        host = {k: v.cpu().pin_memory() for k, v in data.items()}
        h_flat = flat.cpu().pin_memory()
        h_grad = torch.empty_like(h_flat).pin_memory()
        h_tot = torch.zeros(2, dtype=torch.float64).pin_memory()
        hp = lambda k: host[k].data_ptr() if k in host else None  # noqa: E731
        h_st = K.Statics(hp("holding_costs"), hp("underage_costs"), hp("lead_times"), hp("warehouse_lead_times"),
                         hp("warehouse_holding_costs"), hp("warehouse_edge_costs"), hp("echelon_lead_times"),
                         hp("echelon_holding_costs"), hp("mean"), hp("std"))
        h_init = K.State(hp("initial_inventories"), hp("initial_warehouse_inventories"),
                         hp("initial_echelon_inventories"))
        stream = EN.current_stream_ptr(dev)
        h_adj = None
        if pspec.adjacency is not None:
            h_adj = torch.tensor(pspec.adjacency, dtype=torch.int32).contiguous().pin_memory()
        # demand model of the workload for the on-device sampler (workloads.py: per-store mean / std, rho = 0.5, clip)
        if "mean" in host:
            h_mean, h_std, dist_name, rho = host["mean"][0].clone().pin_memory(), host["std"][0].clone().pin_memory(), "normal", 0.5
        elif args.workload == "one_store_lost":
            h_mean, h_std, dist_name, rho = torch.full((1,), 5.0).pin_memory(), torch.ones(1).pin_memory(), "poisson", 0.0
        else:
            sd = 1.6 if args.workload == "one_store_backlogged_lead20" else 2.0
            h_mean, h_std, dist_name, rho = torch.full((1,), 5.0).pin_memory(), torch.full((1,), sd).pin_memory(), "normal", 0.0

        def run_host(generated):
            desc = type(eng.desc).from_buffer_copy(eng.desc)
            step_no = [0]
            if generated:
                desc.demand_source = K.DEMAND_PHILOX_NORMAL if dist_name == "normal" else K.DEMAND_PHILOX_POISSON
                desc.demand_clip_at_zero = 1
                desc.demand_rho = rho
                desc.philox_seed = 57 + rank
                desc.demand_mean = h_mean.data_ptr()
                desc.demand_std = h_std.data_ptr()
                desc.demand_layout = K.DEMAND_TSB
                desc.t_stride = T
            ws_bytes = int(lib.hdpo_rollout_host_workspace_bytes(C.byref(desc)))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            per_step_counters = (T * S * B + 3) // 4

            def host_step():
                if generated:
                    desc.philox_offset = step_no[0] * per_step_counters  # fresh draws every step
                    step_no[0] += 1
                rc = lib.hdpo_rollout_train_host(C.byref(desc), h_flat.data_ptr(),
                                                 None if generated else host["demands"].data_ptr(),
                                                 C.byref(h_st), C.byref(h_init),
                                                 h_adj.data_ptr() if h_adj is not None else None, h_tot.data_ptr(),
                                                 h_grad.data_ptr(), ws.data_ptr(), ws_bytes, stream)
                K.check(lib, rc, "hdpo_rollout_train_host")

            for _ in range(2):
                host_step()
            n_e = max(3, min(args.steps, 5))
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n_e):
                host_step()  # synchronises the stream before returning
            dt = (time.perf_counter() - t0) / n_e
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            h2d = sum(v.numel() * 4 for k, v in host.items() if not (generated and k == "demands")) + h_flat.numel() * 4
            if generated:
                h2d += 8 * S
            del ws
            torch.cuda.empty_cache()
            return {"value": world * B * T / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 16 + h_grad.numel() * 4, "ms_per_step": dt * 1e3,
                    "api": "hdpo_rollout_train_host (C ABI, pinned host buffers)",
                    "demand": ("generated on the device inside the call (Philox4x32-10, " + dist_name + ", fresh draws per step)")
                    if generated else "host tensor uploaded every step"}

        e2e = run_host(True)
        e2e_host_demand = run_host(False)

    # ---- the other BASELINE.json configs, device-timed the same way on a short run (context for the headline line)
    others = None
    if world == 1 and not args.no_others and args.workload == DEFAULT_WORKLOAD and not args.batch:
        del eng
        torch.cuda.empty_cache()
        others = {}
        for name in ("one_warehouse_lost_demand_symmetry_aware", "one_store_backlogged_lead20", "serial_system",
                     "one_store_lost", "many_warehouses_lost_demand", "many_warehouses_lost_demand_8192",
                     "one_store_backlogged_lead20_ckpt10", "one_warehouse_lost_demand_tf32"):
            kw2, ckpt2, prec2 = {}, 0, "tf32x3"
            if name.endswith("_ckpt10"):  # recomputation checkpoints: state of every 10th period taped, rest re-run
                name_wl, ckpt2 = name[:-len("_ckpt10")], 10
            elif name.endswith("_tf32"):  # the headline workload with single-pass TF32 GEMMs (NOT parity grade)
                name_wl, prec2 = name[:-len("_tf32")], "tf32"
            elif name.endswith("_8192"):  # BASELINE cfg 5 in full on ONE GPU: the strong-scaling reference point
                name_wl, kw2 = "many_warehouses_lost_demand", {"B": 8192}
            else:
                name_wl = name
            ps2, pp2, data2, widths2 = WL.WORKLOADS[name_wl](dev, seed=57, T=T, **kw2)
            B2, S2 = data2["demands"].shape[0], pp2["n_stores"]
            flat2 = WL.init_params(widths2, torch.Generator(device=dev).manual_seed(0), dev)
            eng2 = EN.FusedRollout(ps2, pp2, data2, T, ignore_periods=30, precision=prec2, checkpoint_interval=ckpt2)
            grad2 = torch.zeros_like(flat2)

            def step2():
                eng2.forward(flat2, data2)
                eng2.backward(1.0 / (B2 * T * S2), 0.0, out=grad2)

            for _ in range(3):
                step2()
            ms2 = time_region(step2, 10)
            others[name] = {"value": B2 * T / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "scenarios": B2,
                            "precision": prec2, "steps": 10, "warmup": 3, "workspace_bytes": eng2.ws_bytes}
            # its own roofline position: algorithmic FLOPs of the whole step (SymmetryAware: the factored
            # count) against the measured tf32 peak (bf16 / 2; sustained figure: the kernels run inside a long step)
            fl2 = WL.flops_per_scenario_period(widths2, first_layer_dgrad=ps2.arch != "vanilla_serial", n_stores=S2)
            tf2 = fl2 * B2 * T / (ms2 * 1e-3) / 1e12
            pk2 = load_peaks()
            pk2 = pk2.get("bf16_tflops_sustained", pk2["bf16_tflops"]) / 2.0
            others[name]["roofline"] = {"bound": "tensor", "flops_per_scenario_period": fl2, "achieved": tf2,
                                        "peak": pk2, "unit": "TFLOP/s", "step_frac": tf2 / pk2}
            tr2 = load_traffic(name_wl)  # dominant kernel + its dram bytes per launch from the committed ncu capture
            if tr2:
                others[name]["roofline"].update({"kernel": tr2["kernel"], "traffic": tr2["dram_bytes_per_launch"],
                                                 "traffic_source": tr2["source"]})
            if ckpt2:
                others[name]["checkpoint_interval"] = ckpt2
            del eng2, data2, flat2, grad2
            torch.cuda.empty_cache()

    # ---- N > 1: data-parallel equivalence check and the strong-scaling configuration of BASELINE cfg 5
    dp_check, strong = None, None
    if world > 1:
        del eng
        torch.cuda.empty_cache()
        # (a) gradient of ONE global batch sharded over the ranks, after the all-reduce, against the same batch on
        #     rank 0 alone (same seed on every rank -> identical tensors; contiguous shards as parallel.shard_range)
        from neural_inventory_control_b200 import parallel as PL
        Bg = 256 * world
        ps_c, pp_c, data_c, widths_c = WL.WORKLOADS[args.workload](dev, seed=1234, B=Bg, T=T)
        a, b = PL.shard_range(Bg, rank, world)
        shard = {k: v[a:b].contiguous() for k, v in data_c.items()}
        gs = 1.0 / (Bg * T * S)
        e_s = EN.FusedRollout(ps_c, pp_c, shard, T, ignore_periods=30, precision=precision)
        e_s.forward(flat, shard)
        g_sh = e_s.backward(gs, 0.0).clone()
        dist.all_reduce(g_sh)
        if rank == 0:
            e_f = EN.FusedRollout(ps_c, pp_c, data_c, T, ignore_periods=30, precision=precision)
            e_f.forward(flat, data_c)
            g_full = e_f.backward(gs, 0.0)
            rel = float((g_sh.double() - g_full.double()).norm() / g_full.double().norm())
            dp_check = {"scenarios": Bg, "ranks": world, "grad_rel_l2_sharded_vs_single": rel, "bar": 1e-5,
                        "ok": rel <= 1e-5}
            del e_f
        del e_s, data_c, shard
        torch.cuda.empty_cache()
        # (b) strong scaling: many_warehouses_lost_demand, 8192 scenarios GLOBAL (BASELINE cfg 5: "8192 over 8 GPUs")
        Bs = 8192 // world
        ps2, pp2, data2, widths2 = WL.WORKLOADS["many_warehouses_lost_demand"](dev, seed=57 + rank, B=Bs, T=T)
        flat2 = WL.init_params(widths2, torch.Generator(device=dev).manual_seed(0), dev)
        eng2 = EN.FusedRollout(ps2, pp2, data2, T, ignore_periods=30, precision=precision)
        grad2 = torch.zeros_like(flat2)
        S2 = pp2["n_stores"]

        def step_s():
            eng2.forward(flat2, data2)
            eng2.backward(1.0 / (8192 * T * S2), 0.0, out=grad2)
            dist.all_reduce(grad2)

        for _ in range(3):
            step_s()
        dist.barrier()
        ms_s = time_region(step_s, 8)
        t = torch.tensor([ms_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_s = float(t.item())
        strong = {"workload": "many_warehouses_lost_demand", "scaling": "strong", "scenarios_global": 8192,
                  "scenarios_per_gpu": Bs, "ms_per_step": ms_s, "value": 8192 * T / (ms_s * 1e-3), "unit": UNIT,
                  "note": "single-GPU time of the same 8192-scenario batch: other_workloads."
                          "many_warehouses_lost_demand_8192 of the N = 1 line"}
        del eng2, data2

    cpu_baseline = None
    reference_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = cpu_sample_size(args.workload)
        n_cpu = 3 if Bc * S >= 16384 else 10
        v, dt, cores, kind, note = time_reference(args.workload, Bc, T, n_cpu, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": dt * 1e3,
                        "sample": f"{Bc} scenarios x {T} periods, 1 warm-up + {n_cpu} timed steps, {note}, {cores} host threads"}
        # the same unmodified reference code with device='cuda:0' on this B200 (stock PyTorch eager kernels): the
        # same-box number SURVEY.md 8d asks for, on the SAME scenarios-per-step as our arm where it fits
        try:
            torch.cuda.empty_cache()
            Bg = min(B, 8192)
            vg, dtg, _, kindg, noteg = time_reference(args.workload, Bg, T, 3, 1, device=f"cuda:{local}")
            reference_cuda = {"value": vg, "unit": UNIT, "ms_per_step": dtg * 1e3, "scenarios": Bg, "kind": kindg,
                              "note": f"{noteg}, device cuda:{local}, fp32 (allow_tf32 off), 1 warm-up + 3 timed steps"}
        except Exception as e:  # noqa: BLE001  (no such policy in the reference snapshot, or out of memory)
            reference_cuda = {"unavailable": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else ("tf32x3 (fp32-grade split, fp32 accumulate)" if precision == "tf32x3"
                                                         else "tf32"),
            "data": "synthetic (value: demand resident in HBM; e2e: demand generated on the device by the Philox sampler)",
            "config": {"workload": args.workload, "scenarios_per_gpu": B, "periods": T, "stores": S,
                       "policy_widths": widths, "l2": "inputs larger than L2 (demand + state tape per step)"
                       if B * T * 4 * (1 + WL.net_list(widths)[0][1][0]) > 126e6 else "working set below L2 size; no flush",
                       "parallelism": f"dp{world} (scenario shards, gradient all-reduce)" if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "e2e_host_demand": e2e_host_demand, "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline, "reference_cuda": reference_cuda, "other_workloads": others,
            "dp_check": dp_check, "strong_scaling": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
