"""N > 1 host logic on CPU: world_size-2 gloo process group, scenario shards, one bucketed gradient all-reduce.

The compute inside each rank is the PyTorch-eager port from oracle/ (tests may use the oracle); what is under test
is the sharding / reduction logic of neural_inventory_control_b200.parallel that bench.py and the Trainer use on
the GPUs: the sum of shard gradients (each scaled by 1/(B_global*T*S)) must equal the single-process gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _policy(meta, params):
    idxs = sorted({int(k.split(".")[2]) for k in params if k.endswith(".weight")})
    layers = [(torch.tensor(params[f"net.master.{i}.weight"], requires_grad=True),
               torch.tensor(params[f"net.master.{i}.bias"], requires_grad=True)) for i in idxs]
    return {"arch": meta["nn_name"], "layers": layers, "hidden_act": meta["inner_layer_activations"]["master"],
            "out_act": meta["output_layer_activation"]["master"],
            "wub": torch.tensor([meta["warehouse_upper_bound"]]), "adjacency": None, "transshipment": False}


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from neural_inventory_control_b200 import parallel as PL
    from oracle import torch_port as TP
    r, w, _ = PL.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) == PL.world_info()
    meta, g = G.load("rollout", name)
    data = {k: torch.tensor(v) for k, v in g["data"].items()}
    B, T = data["demands"].shape[0], 12
    pb = dict(meta["problem_params"], period_shift=0)
    shard = PL.shard_batch(data, rank, world)
    pol = _policy(meta, g["param"])
    total, report, _ = TP.simulate(pol, pb, shard, T, 4)
    (total / (B * T * pb["n_stores"])).backward()   # GLOBAL batch size in the scale

    class M:  # minimal stand-in for a module: parameters() with .grad
        def parameters(self_inner):
            return [t for wb in pol["layers"] for t in wb]
    tot, rep = total.detach().clone(), report.detach().clone()
    PL.allreduce_gradients_and_losses(M(), [tot, rep])
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), total=tot.numpy(), report=rep.numpy(),
                 **{f"g{i}": t.grad.numpy() for i, t in enumerate(M().parameters())})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from neural_inventory_control_b200 import parallel as PL
    for n in (0, 1, 7, 8, 64, 1000003):
        for world in (1, 2, 3, 8):
            edges = [PL.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("name", ["one_store_lost", "serial_system"])
def test_two_rank_gradient_equals_single_process(tmp_path, name):
    from oracle import torch_port as TP
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "reduced.npz"))
    meta, g = G.load("rollout", name)
    data = {k: torch.tensor(v) for k, v in g["data"].items()}
    B, T = data["demands"].shape[0], 12
    pb = dict(meta["problem_params"], period_shift=0)
    pol = _policy(meta, g["param"])
    total, report, _ = TP.simulate(pol, pb, data, T, 4)
    (total / (B * T * pb["n_stores"])).backward()
    assert abs(float(got["total"]) / float(total) - 1) < 1e-6
    assert abs(float(got["report"]) / float(report) - 1) < 1e-6
    for i, t in enumerate(t for wb in pol["layers"] for t in wb):
        assert G.rel_l2(got[f"g{i}"], t.grad.numpy()) < 1e-6   # summation order differs across shards
