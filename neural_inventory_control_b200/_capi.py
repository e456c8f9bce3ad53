"""ctypes mirror of include/hdpo_b200.h (struct layouts + prototypes).

`bind(cdll)` attaches argtypes/restypes to a loaded library. The product loader (`_lib.py`) binds
libhdpo_b200.so; the CPU tests bind the host-thread emulation build of the same sources.
Pointers are passed as integer addresses (`tensor.data_ptr()` / `ndarray.ctypes.data`).
"""
import ctypes as C

HDPO_MAX_LAYERS = 8

ARCH = {"vanilla_one_store": 0, "vanilla_serial": 1, "vanilla_warehouse": 2, "symmetry_aware": 3}
ACT = {None: 0, "elu": 1, "relu": 2, "tanh": 3, "sigmoid": 4, "softplus": 5}
DEMAND_BST, DEMAND_TSB = 0, 1
PREC = {"fp32": 0, "tf32x3": 1, "tf32": 2}

p = C.c_void_p


class Problem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "S", "W", "E", "L", "Lw", "Le", "lost_demand", "maximize_profit", "has_edge_cost")]


class Statics(C.Structure):
    _fields_ = [(n, p) for n in
                ("holding_costs", "underage_costs", "lead_times", "warehouse_lead_times", "warehouse_holding_costs",
                 "warehouse_edge_costs", "echelon_lead_times", "echelon_holding_costs", "mean", "std")]


class State(C.Structure):
    _fields_ = [("store", p), ("warehouse", p), ("echelon", p)]


class Action(C.Structure):
    _fields_ = [("stores", p), ("warehouses", p), ("echelons", p)]


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("widths", C.c_int32 * (HDPO_MAX_LAYERS + 1)), ("hidden_act", C.c_int32),
                ("out_act", C.c_int32)]


class RolloutDesc(C.Structure):
    _fields_ = [("pb", Problem)] + [(n, C.c_int32) for n in
                                    ("arch", "T", "t_stride", "period_shift", "ignore_periods", "demand_layout",
                                     "discrete_allocation", "transshipment", "precision", "save_for_backward")] + \
               [("warehouse_upper_bound", C.c_float), ("prop_eps", C.c_float), ("master", Mlp), ("store_net", Mlp),
                ("warehouse_net", Mlp), ("adjacency", p),
                # ABI 2: demand generated on the device by the Philox sampler (K4 in the path)
                ("demand_source", C.c_int32), ("demand_clip_at_zero", C.c_int32), ("demand_rho", C.c_float),
                ("checkpoint_interval", C.c_int32), ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
                ("demand_mean", p), ("demand_std", p)]


DEMAND_FROM_ARGUMENT, DEMAND_PHILOX_NORMAL, DEMAND_PHILOX_POISSON = 0, 1, 2


class HdpoError(RuntimeError):
    pass


def make_mlp(widths, hidden_act, out_act):
    """Nets the kernels do not cover (more than HDPO_MAX_LAYERS layers, activations outside ACT such as the
    reference's 'softmax') raise the "no fused rollout" error that Trainer.simulate_batch turns into the generic
    per-step path, exactly like an unsupported policy shape."""
    m = Mlp()
    if len(widths) - 1 > HDPO_MAX_LAYERS:
        raise HdpoError(f"no fused rollout: MLP with {len(widths) - 1} linear layers exceeds HDPO_MAX_LAYERS={HDPO_MAX_LAYERS}")
    if hidden_act not in ACT or out_act not in ACT:
        raise HdpoError(f"no fused rollout: activation {hidden_act!r} / {out_act!r} is not implemented in the kernels")
    m.n_layers = len(widths) - 1
    for i, w in enumerate(widths):
        m.widths[i] = int(w)
    m.hidden_act = ACT[hidden_act]
    m.out_act = ACT[out_act]
    return m


EXPORTS = [
    "hdpo_step_fwd", "hdpo_step_bwd", "hdpo_allocation_shift", "hdpo_gather_rows", "hdpo_adam_step", "hdpo_param_count", "hdpo_rollout_workspace_bytes",
    "hdpo_rollout_fwd", "hdpo_rollout_bwd", "hdpo_rollout_host_workspace_bytes", "hdpo_rollout_train_host",
    "hdpo_philox_normal", "hdpo_philox_poisson", "hdpo_philox_raw", "hdpo_debug_gemm_tc", "hdpo_debug_gemm_tc_wgrad", "hdpo_debug_gemm_tc_timeline", "hdpo_debug_set_trace", "hdpo_debug_set_wp_trace", "hdpo_debug_set_wide_persist", "hdpo_debug_set_tc_multi", "hdpo_debug_set_tc_occ2", "hdpo_debug_set_wide_wg_overlap", "hdpo_debug_set_wide_ksplit", "hdpo_debug_set_small_unit", "hdpo_debug_set_small_unit_group", "hdpo_last_error", "hdpo_abi_version", "hdpo_kernel_launch_count",
    "hdpo_device_info",
]


def bind(lib):
    P = C.POINTER
    lib.hdpo_step_fwd.argtypes = [P(Problem), P(Statics), P(State), P(Action), p, C.c_int64, C.c_int64, P(State), p, p]
    lib.hdpo_step_bwd.argtypes = [P(Problem), P(Statics), P(State), P(Action), p, C.c_int64, C.c_int64, P(State), p,
                                  P(State), P(Action), p]
    lib.hdpo_allocation_shift.argtypes = [p, C.c_int32, C.c_int32, C.c_int32, p]
    lib.hdpo_gather_rows.argtypes = [p, p, p, C.c_int64, C.c_int64, p]
    lib.hdpo_adam_step.argtypes = [p, p, p, p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, p]
    lib.hdpo_param_count.argtypes = [P(RolloutDesc)]
    lib.hdpo_param_count.restype = C.c_int64
    lib.hdpo_rollout_workspace_bytes.argtypes = [P(RolloutDesc)]
    lib.hdpo_rollout_workspace_bytes.restype = C.c_size_t
    lib.hdpo_rollout_fwd.argtypes = [P(RolloutDesc), p, p, P(Statics), P(State), p, p, p, p, P(State), p, C.c_size_t, p]
    lib.hdpo_rollout_bwd.argtypes = [P(RolloutDesc), p, p, P(Statics), C.c_float, C.c_float, p, p, C.c_size_t, p]
    lib.hdpo_rollout_host_workspace_bytes.argtypes = [P(RolloutDesc)]
    lib.hdpo_rollout_host_workspace_bytes.restype = C.c_size_t
    lib.hdpo_rollout_train_host.argtypes = [P(RolloutDesc), p, p, P(Statics), P(State), p, p, p, p, C.c_size_t, p]
    lib.hdpo_philox_normal.argtypes = [p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, p, p, C.c_float, C.c_int32,
                                       C.c_uint64, C.c_uint64, p]
    lib.hdpo_philox_poisson.argtypes = [p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, p, C.c_uint64, C.c_uint64, p]
    lib.hdpo_philox_raw.argtypes = [p, C.c_uint64, C.c_uint64, C.c_uint64, p]
    lib.hdpo_debug_gemm_tc.argtypes = [p, p, p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, p, p]
    lib.hdpo_debug_gemm_tc_wgrad.argtypes = [p, p, p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, p, p]
    lib.hdpo_debug_gemm_tc_timeline.argtypes = [p, p, p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, p, p, p, C.c_int32, p, p]
    lib.hdpo_debug_set_trace.argtypes = [p, C.c_int64]
    lib.hdpo_debug_set_wp_trace.argtypes = [p, C.c_int32]
    lib.hdpo_debug_set_wide_persist.argtypes = [C.c_int32]
    lib.hdpo_debug_set_tc_multi.argtypes = [C.c_int32]
    lib.hdpo_debug_set_tc_occ2.argtypes = [C.c_int32]
    lib.hdpo_debug_set_wide_wg_overlap.argtypes = [C.c_int32, C.c_int32]
    lib.hdpo_debug_set_wide_ksplit.argtypes = [C.c_int32]
    lib.hdpo_debug_set_small_unit.argtypes = [C.c_int32]
    lib.hdpo_debug_set_small_unit_group.argtypes = [C.c_int32]
    lib.hdpo_last_error.restype = C.c_char_p
    lib.hdpo_abi_version.restype = C.c_int
    lib.hdpo_kernel_launch_count.restype = C.c_int64
    lib.hdpo_device_info.argtypes = [P(C.c_int32), P(C.c_int32), P(C.c_int32), C.c_char_p, C.c_int32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:  # default restype: int status
            pass
    return lib


def check(lib, rc, what):
    if rc != 0:
        raise HdpoError(f"{what} failed (code {rc}): {lib.hdpo_last_error().decode(errors='replace')}")
