"""Scenario-parallel training over the GPUs of one box: one process per GPU (torchrun), NCCL over NVLink.

The rollout path shards perfectly: scenarios are independent, the only coupling is the policy gradient
(SURVEY.md 8e). Each rank runs the fused forward + adjoint kernels on its contiguous shard of every batch with
dLoss/dtotal = 1/(B_global*T*S), then ONE bucketed all-reduce (sum) carries the flat gradient plus the two loss
scalars; the optimizer step is replicated. Payloads are tiny (9 KB .. 3 MB), i.e. latency-bound: a single
collective per batch, no per-layer buckets.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world, local_rank).
    HDPO_DIST_BACKEND overrides the backend (tests run two ranks on ONE GPU over gloo)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or os.environ.get("HDPO_DIST_BACKEND")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world):
    """Contiguous shard [start, stop) of n scenarios for `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(data, rank, world):
    """Views of the rank's scenarios of every tensor in a batch dict (dim 0 = scenario)."""
    if world == 1:
        return data
    n = len(data["initial_inventories"])
    a, b = shard_range(n, rank, world)
    return {k: v[a:b] for k, v in data.items()}


def allreduce_sum_(tensors):
    """In-place sum over ranks of a list of tensors, as ONE flat bucket (one collective per batch)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n
    return tensors


def broadcast_parameters_(model, src=0):
    """Rank `src`'s parameters to every rank as ONE flat bucket (replicas must start from identical weights; the
    LazyLinear layers of the reference's nets materialise at the first batch, so this runs right after it)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    params = [p for p in model.parameters() if not isinstance(p, torch.nn.parameter.UninitializedParameter)]
    if not params:
        return
    flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])
    dist.broadcast(flat, src=src)
    o = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[o:o + n].view_as(p))
            o += n


def allreduce_gradients_and_losses(model, losses):
    """Sum parameter gradients and the given 0-d loss tensors over ranks with a single collective."""
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    allreduce_sum_(grads + list(losses))
