"""Pointer-agnostic construction of the C-ABI descriptors (no torch import here).

Shared by the product host code (torch tensors -> data_ptr) and by the CPU tests that drive the
host-thread emulation build with numpy arrays.
"""
from . import _capi as K


def problem(B, S, W, E, L, Lw=0, Le=0, lost_demand=True, maximize_profit=False, has_edge_cost=False):
    pb = K.Problem()
    pb.B, pb.S, pb.W, pb.E = int(B), int(S), int(W), int(E)
    pb.L, pb.Lw, pb.Le = int(L), int(Lw), int(Le)
    pb.lost_demand, pb.maximize_profit, pb.has_edge_cost = int(bool(lost_demand)), int(bool(maximize_profit)), int(
        bool(has_edge_cost))
    return pb


def mlp_widths(state_shapes):
    """widths from an ordered list of weight shapes [(out,in), ...]."""
    widths = [state_shapes[0][1]]
    for out, inp in state_shapes:
        if inp != widths[-1]:
            raise ValueError(f"layer input {inp} does not chain with previous width {widths[-1]}")
        widths.append(out)
    return widths


def rollout_desc(arch, pb, T, t_stride, master, period_shift=0, ignore_periods=0, demand_layout=K.DEMAND_BST,
                 discrete_allocation=False, transshipment=False, precision="fp32", save_for_backward=True,
                 warehouse_upper_bound=0.0, prop_eps=1e-15, store_net=None, warehouse_net=None, adjacency_ptr=None,
                 philox=None, checkpoint_interval=0):
    """master / store_net / warehouse_net: (widths, hidden_act, out_act).
    philox: None (demand comes from the `demands` argument) or a dict {dist: 'normal' | 'poisson', mean_ptr, std_ptr,
    rho, clip, seed, offset}: the [t_stride, S, B] trace is generated on the device inside the forward call."""
    d = K.RolloutDesc()
    d.pb = pb
    d.arch = K.ARCH[arch]
    d.T, d.t_stride, d.period_shift, d.ignore_periods = int(T), int(t_stride), int(period_shift), int(ignore_periods)
    d.demand_layout = int(demand_layout)
    d.discrete_allocation = int(bool(discrete_allocation))
    d.transshipment = int(bool(transshipment))
    d.precision = K.PREC[precision]
    d.save_for_backward = int(bool(save_for_backward))
    d.warehouse_upper_bound = float(warehouse_upper_bound)
    d.prop_eps = float(prop_eps)
    d.master = K.make_mlp(*master)
    if store_net is not None:
        d.store_net = K.make_mlp(*store_net)
    if warehouse_net is not None:
        d.warehouse_net = K.make_mlp(*warehouse_net)
    d.adjacency = adjacency_ptr
    d.checkpoint_interval = int(checkpoint_interval)
    if philox is not None:
        d.demand_source = K.DEMAND_PHILOX_NORMAL if philox["dist"] == "normal" else K.DEMAND_PHILOX_POISSON
        d.demand_clip_at_zero = int(bool(philox.get("clip", False)))
        d.demand_rho = float(philox.get("rho", 0.0))
        d.philox_seed = int(philox.get("seed", 0))
        d.philox_offset = int(philox.get("offset", 0))
        d.demand_mean = philox["mean_ptr"]
        d.demand_std = philox.get("std_ptr")
        d.demand_layout = K.DEMAND_TSB
    return d
