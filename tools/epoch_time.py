"""Wall-clock of full training epochs through the reference-facing API (Trainer.train) on the GPU box.

    python tools/epoch_time.py [setting] [policy] [epochs]

Reports seconds per epoch with the device-resident dataset (default) and with the plain torch DataLoader path
(HDPO_DEVICE_DATASET=0 semantics), i.e. what `python main_run.py train <setting> <policy>` costs per epoch."""
import copy
import os
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import yaml  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario  # noqa: E402
from neural_inventory_control_b200.environment import Simulator  # noqa: E402
from neural_inventory_control_b200.loss_functions import PolicyLoss  # noqa: E402
from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator  # noqa: E402
from neural_inventory_control_b200.trainer import Trainer  # noqa: E402


def main():
    setting = sys.argv[1] if len(sys.argv) > 1 else "one_store_lost"
    policy = sys.argv[2] if len(sys.argv) > 2 else "vanilla_one_store"
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    s = yaml.safe_load(open(f"{ROOT}/config_files/settings/{setting}.yml"))
    p = yaml.safe_load(open(f"{ROOT}/config_files/policies_and_hyperparams/{policy}.yml"))
    obs = defaultdict(lambda: None, s["observation_params"])
    pbd = s["params_by_dataset"]
    common = (s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"])
    t0 = time.time()
    sc = Scenario(max(pbd["train"]["periods"], pbd["dev"]["periods"]), *common,
                  pbd["train"]["n_samples"] + pbd["dev"]["n_samples"], obs, copy.deepcopy(s["seeds"]))
    train, dev = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True,
                                                  sample_index_for_split=pbd["dev"]["n_samples"])
    print(f"scenario generation: {time.time() - t0:.2f} s")
    devc = "cuda:0"
    for use_dd in (True, False):
        loaders = {"train": DataLoader(train, batch_size=pbd["train"]["batch_size"], shuffle=True),
                   "dev": DataLoader(dev, batch_size=pbd["dev"]["batch_size"], shuffle=False)}
        torch.manual_seed(0)
        model = NeuralNetworkCreator().create_neural_network(sc, p["nn_params"], device=devc)
        opt = torch.optim.Adam(model.parameters(), lr=p["optimizer_params"]["learning_rate"])
        tr, sim = Trainer(device=devc), Simulator(device=devc)
        tr.use_device_dataset = use_dd
        tp = dict(p["trainer_params"], epochs=epochs, do_dev_every_n_epochs=10 ** 9, print_results_every_n_epochs=10 ** 9,
                  save_model=False)
        tr.train(1, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs, pbd, tp)  # warm-up epoch
        torch.cuda.synchronize()
        t0 = time.time()
        tr.train(epochs, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs, pbd, tp)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / epochs
        n = pbd["train"]["n_samples"] * pbd["train"]["periods"]
        print(f"{setting}/{policy} device_dataset={use_dd}: {dt * 1e3:.1f} ms per epoch "
              f"({n / dt / 1e6:.1f} M scenario-periods/s incl. batch assembly + Adam), path={tr.last_path}, "
              f"train loss {tr.all_train_losses[-1]:.4f}")


if __name__ == "__main__":
    main()
