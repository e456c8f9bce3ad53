"""cProfile of Trainer.train epochs on the GPU box (host-side overhead hunting)."""
import cProfile, pstats, sys, os, copy, io
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from torch.utils.data import DataLoader
from neural_inventory_control_b200.data_handling import DatasetCreator, Scenario
from neural_inventory_control_b200.environment import Simulator
from neural_inventory_control_b200.loss_functions import PolicyLoss
from neural_inventory_control_b200.neural_networks import NeuralNetworkCreator
from neural_inventory_control_b200.trainer import Trainer
setting, policy = "one_store_lost", "vanilla_one_store"
s = yaml.safe_load(open(f"{ROOT}/config_files/settings/{setting}.yml"))
p = yaml.safe_load(open(f"{ROOT}/config_files/policies_and_hyperparams/{policy}.yml"))
obs = defaultdict(lambda: None, s["observation_params"]); pbd = s["params_by_dataset"]
common = (s["problem_params"], s["store_params"], s["warehouse_params"], s["echelon_params"])
sc = Scenario(100, *common, 65536, obs, copy.deepcopy(s["seeds"]))
train, dev = DatasetCreator().create_datasets(sc, split=True, by_sample_indexes=True, sample_index_for_split=32768)
loaders = {"train": DataLoader(train, batch_size=8192, shuffle=True), "dev": DataLoader(dev, batch_size=32768)}
model = NeuralNetworkCreator().create_neural_network(sc, p["nn_params"], device="cuda:0")
opt = torch.optim.Adam(model.parameters(), lr=0.003)
tr, sim = Trainer(device="cuda:0"), Simulator(device="cuda:0")
tp = dict(p["trainer_params"], do_dev_every_n_epochs=10 ** 9, print_results_every_n_epochs=10 ** 9, save_model=False)
tr.train(2, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs, pbd, tp)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
tr.train(10, PolicyLoss(), sim, model, loaders, opt, s["problem_params"], obs, pbd, tp)
torch.cuda.synchronize(); pr.disable()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(35); print(out.getvalue()[:6000])
