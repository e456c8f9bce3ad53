"""`python main_run.py <train|test> [<setting> <policy>]` - same command line and wiring as the reference's
main_run.py (argv handling, YAML keys, scenario/dataset/model/optimizer construction, train-or-test dispatch),
running on the B200-native engine: fusable policies train through the fused forward/adjoint rollout kernels.
"""
import sys

import yaml

from trainer import *  # noqa: F401,F403  (star-import chain of the reference: torch, DefaultDict, DataLoader, ...)

SETTING_KEYS = ("seeds", "test_seeds", "problem_params", "params_by_dataset", "observation_params", "store_params",
                "warehouse_params", "echelon_params", "sample_data_params")
HYPERPARAM_KEYS = ("trainer_params", "optimizer_params", "nn_params")


def parse_argv(argv):
    if len(argv) == 4:
        return argv[1], argv[2], argv[3]
    if len(argv) == 2:
        return argv[1], "one_store_lost", "vanilla_one_store"
    print(f"Number of parameters provided including script name: {len(argv)}")
    print("Number of parameters should be either 4 or 2 (so that last 2 parameters defined in main_run.py)")
    raise SystemExit(1)


def load_yaml(path):
    with open(path, "r") as f:
        return yaml.safe_load(f)


def build_datasets(cfg, observation_params):
    """Train/dev/test datasets + the scenario the model is built from (the TEST scenario for synthetic settings,
    exactly as the reference rebinds it)."""
    pbd, creator = cfg["params_by_dataset"], DatasetCreator()
    common = (cfg["problem_params"], cfg["store_params"], cfg["warehouse_params"], cfg["echelon_params"])
    if cfg["sample_data_params"]["split_by_period"]:
        scenario = Scenario(None, *common, pbd["train"]["n_samples"], observation_params, cfg["seeds"])
        windows = [cfg["sample_data_params"][k] for k in ("train_periods", "dev_periods", "test_periods")]
        train, dev, test = creator.create_datasets(scenario, split=True, by_period=True, periods_for_split=windows)
        return scenario, train, dev, test
    periods = max(pbd["train"]["periods"], pbd["dev"]["periods"])
    scenario = Scenario(periods, *common, pbd["train"]["n_samples"] + pbd["dev"]["n_samples"], observation_params,
                        cfg["seeds"])
    train, dev = creator.create_datasets(scenario, split=True, by_sample_indexes=True,
                                         sample_index_for_split=pbd["dev"]["n_samples"])
    scenario = Scenario(pbd["test"]["periods"], *common, pbd["test"]["n_samples"], observation_params, cfg["test_seeds"])
    test = creator.create_datasets(scenario, split=False)
    return scenario, train, dev, test


def main(argv=None):
    mode, setting_name, hyperparams_name = parse_argv(sys.argv if argv is None else argv)
    print(f"Setting file name: {setting_name}")
    print(f"Hyperparams file name: {hyperparams_name}\n")
    setting = load_yaml(f"config_files/settings/{setting_name}.yml")
    hyper = load_yaml(f"config_files/policies_and_hyperparams/{hyperparams_name}.yml")
    cfg = {k: setting[k] for k in SETTING_KEYS}
    trainer_params, optimizer_params, nn_params = (hyper[k] for k in HYPERPARAM_KEYS)
    observation_params = DefaultDict(lambda: None, cfg["observation_params"])
    problem_params, pbd = cfg["problem_params"], cfg["params_by_dataset"]

    # one process per GPU under torchrun (WORLD_SIZE > 1): scenario shards of every batch, ONE gradient all-reduce per
    # batch over NCCL (neural_inventory_control_b200/parallel.py); plain `python main_run.py ...` is rank 0 of 1
    from neural_inventory_control_b200 import parallel as PL
    rank, world, local_rank = PL.init_from_env()
    device = f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu"
    if rank != 0:  # every rank trains; rank 0 alone talks and writes checkpoints
        import builtins
        builtins.print = lambda *a, **k: None
        trainer_params["save_model"] = False
    scenario, train_set, dev_set, test_set = build_datasets(cfg, observation_params)
    data_loaders = {
        "train": DataLoader(train_set, batch_size=pbd["train"]["batch_size"], shuffle=True),
        "dev": DataLoader(dev_set, batch_size=pbd["dev"]["batch_size"], shuffle=False),
        "test": DataLoader(test_set, batch_size=pbd["test"]["batch_size"], shuffle=False),
    }
    model = NeuralNetworkCreator().create_neural_network(scenario, nn_params, device=device)
    loss_function = PolicyLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=optimizer_params["learning_rate"])
    simulator = Simulator(device=device)
    trainer = Trainer(device=device)

    trainer_params["base_dir"] = "saved_models"
    trainer_params["save_model_folders"] = [trainer.get_year_month_day(), nn_params["name"]]
    trainer_params["save_model_filename"] = trainer.get_time_stamp()
    if trainer_params["load_previous_model"]:
        print(f"Loading model from {trainer_params['load_model_path']}")
        model, optimizer = trainer.load_model(model, optimizer, trainer_params["load_model_path"])

    if mode == "train":  # like the reference, `train` does not run the test set afterwards
        trainer.train(trainer_params["epochs"], loss_function, simulator, model, data_loaders, optimizer,
                      problem_params, observation_params, pbd, trainer_params)
    elif mode == "test":
        _, report = trainer.test(loss_function, simulator, model, data_loaders, optimizer, problem_params,
                                 observation_params, pbd,
                                 discrete_allocation=cfg["store_params"]["demand"]["distribution"] == "poisson")
        print(f"Average per-period test loss: {report}")
    else:
        print(f"Invalid argument: {mode}")
        raise SystemExit(1)


if __name__ == "__main__":
    main()
