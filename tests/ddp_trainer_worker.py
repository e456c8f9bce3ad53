"""Worker of tests/test_trainer_gpu.py::test_two_rank_trainer_matches_single_process (TEST INFRASTRUCTURE).

Trains the shipped one_store_lost / vanilla_one_store setting (reduced sample counts) for a few epochs through
`Trainer.train` exactly as main_run.py wires it, under WORLD_SIZE ranks that share cuda:0 (HDPO_DIST_BACKEND=gloo), and
writes rank 0's final parameters + loss history to argv[1]."""
import copy
import os
import sys
from collections import defaultdict

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    from torch.utils.data import DataLoader
    from neural_inventory_control_b200 import parallel as PL
    from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
    from neural_inventory_control_b200.environment import Simulator
    from neural_inventory_control_b200.loss_functions import PolicyLoss
    from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
    from neural_inventory_control_b200.trainer import Trainer
    rank, world, _ = PL.init_from_env()
    dev = "cuda:0"

    def cfg(kind, name):
        with open(os.path.join(ROOT, "config_files", kind, f"{name}.yml")) as f:
            return yaml.safe_load(f)
    s = copy.deepcopy(cfg("settings", "one_store_lost"))
    p = copy.deepcopy(cfg("policies_and_hyperparams", "vanilla_one_store"))
    obs_params = defaultdict(lambda: None, s["observation_params"])
    pbd = s["params_by_dataset"]
    pbd["train"].update(n_samples=1000, batch_size=300)  # ragged: 4 batches, the last one of 100; 300 / 2 ranks
    pbd["dev"].update(n_samples=256, batch_size=256, periods=60, ignore_periods=30)
    common = (s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"])
    sc = Scenario(60, *common, 1000 + 256, obs_params, s["seeds"])
    train, devset = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True, sample_index_for_split=256)
    loaders = {"train": DataLoader(train, batch_size=300, shuffle=True),
               "dev": DataLoader(devset, batch_size=256, shuffle=False)}
    torch.manual_seed(100 + rank)  # DIFFERENT torch seeds per rank: rank 0's permutation and weights must win
    model = NeuralNetworkCreator().create_neural_network(sc, p["nn_params"], device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=p["optimizer_params"]["learning_rate"])
    tr, sim = Trainer(device=dev), Simulator(device=dev)
    tp = p["trainer_params"]
    tp.update(epochs=4, do_dev_every_n_epochs=2, print_results_every_n_epochs=100, save_model=False)
    tr.train(4, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs_params, pbd, tp)
    assert tr.last_path == "fused"
    if rank == 0:
        np.savez(out_path, train=np.array(tr.all_train_losses), dev=np.array(tr.all_dev_losses),
                 **{k.replace(".", "_"): v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
