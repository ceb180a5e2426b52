"""Entry point logic behind ``e2e_trainer.py`` (ref. ``e2e_trainer.py``: argparse :200-214, output tree
:222-235, logging :238, config :240-250, ``run_worker`` :77-195).

Same flags: ``-config -outputPath -dataPath -task -backend {nccl,gloo} [-num_skip_decoding] [--local_rank]``
launched with ``python -m torch.distributed.run --nproc_per_node=N e2e_trainer.py …`` (or plain ``python`` for
one process).  Differences: ``experiment_name`` is deterministic and identical on every rank
(``-experiment`` flag / ``FLUTE_EXPERIMENT`` env / config file stem) — the reference derives it from a random
per-process AzureML offline run id, which breaks ``resume_from_checkpoint`` outside AzureML;
rendezvous defaults to 127.0.0.1; rank 0's GPU trains clients too.
"""
from __future__ import annotations

import argparse
import logging
import os
import shutil

import torch
import yaml

from .core import federated
from .core.client import Client
from .core.config import FLUTEConfig
from .core.evaluation import make_eval_clients
from .core.server import select_server
from .models import make_model
from .parallel.comm import init_distributed, make_communicator
from .utils import find_pretrained_model, init_logging, make_optimizer, print_rank
from .utils.dataloaders_utils import get_dataset, make_train_dataloader
from .utils.metrics_sink import get_run


def log_run_properties(config):
    from psutil import virtual_memory
    props = {"System memory (GB)": float(virtual_memory().total) / (1024 ** 3)}
    for key, default in [("server_config.num_clients_per_iteration", 0), ("server_config.max_iteration", 0),
                         ("dp_config.eps", 0), ("dp_config.max_weight", 0), ("dp_config.min_weight", 0),
                         ("server_config.optimizer_config.type", "sgd"), ("server_config.optimizer_config.lr", 1.0),
                         ("server_config.optimizer_config.amsgrad", False),
                         ("server_config.annealing_config.type", "step_lr"),
                         ("server_config.annealing_config.step_interval", "epoch"),
                         ("server_config.annealing_config.gamma", 1.0),
                         ("server_config.annealing_config.step_size", 100)]:
        props[key] = config.lookup(key, default)
    run = get_run()
    for k, v in props.items():
        run.log(k, v)


def build_job(config, task, data_path, model_path, backend=None):
    """Construct everything one rank needs: (server_or_None, worker, comm)."""
    import copy
    model_config, server_config, client_config = config["model_config"], config["server_config"], config["client_config"]
    b200 = server_config.get("b200", {}) or {}
    init_distributed(backend)
    comm = make_communicator(b200.get("comm", "auto"))
    seed = int(b200.get("seed", 0))
    import random
    import numpy as np
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)

    model = make_model(model_config)
    val_dataset = get_dataset(data_path, config, task, mode="val", test_only=True)
    test_dataset = get_dataset(data_path, config, task, mode="test", test_only=True)
    val_clients = list(make_eval_clients(val_dataset, config))
    test_clients = list(make_eval_clients(test_dataset, config))
    num_clients = Client.get_train_dataset(data_path, config, task)
    config["server_config"]["data_config"]["num_clients"] = num_clients

    # every rank owns a worker with its OWN model replica (the server model is never trained in place)
    worker_model = copy.deepcopy(model)
    worker = federated.Worker(model=worker_model, data_path=data_path,
                              do_profiling=client_config.get("do_profiling", False), val_clients=val_clients,
                              test_clients=test_clients, val_dataset=val_dataset, test_dataset=test_dataset,
                              config=config)
    federated.init_runtime(comm, worker, server_is_worker=b200.get("server_is_worker", True),
                           dispatch=b200.get("dispatch", "static_lpt"))
    if b200.get("device_engine", True) and torch.cuda.is_available():
        try:
            from .core.engine import DeviceClientEngine
            worker.engine = DeviceClientEngine.maybe_create(worker, config, task)
        except Exception as e:  # the engine is an accelerator, never a requirement
            print_rank("device engine unavailable: {}".format(e), logging.WARNING)

    server = None
    if comm.rank == 0:
        print_rank("Server data preparation")
        server_train_dataloader = None
        if "train" in server_config["data_config"]:
            server_train_dataloader = make_train_dataloader(server_config["data_config"]["train"], data_path,
                                                            task=task, clientx=None)
        optimizer = make_optimizer(server_config["optimizer_config"], model)
        best_trained_model = find_pretrained_model(model_path, model_config)
        if best_trained_model is not None and os.path.exists(best_trained_model):
            sd = torch.load(best_trained_model, map_location=None if torch.cuda.is_available() else "cpu",
                            weights_only=False)
            model.load_state_dict(sd.get("model_state_dict", sd) if isinstance(sd, dict) else sd)
        from .core import client as client_mod
        costs = list(client_mod.train_dataset.num_samples)
        server = select_server(server_config.get("type", "model_optimization"))(
            num_clients=num_clients, model=model, optimizer=optimizer, ss_scheduler=None, data_path=data_path,
            model_path=model_path, server_train_dataloader=server_train_dataloader, config=config,
            idx_val_clients=list(range(len(val_clients))), idx_test_clients=list(range(len(test_clients))),
            single_worker=worker, client_costs=costs)
        log_run_properties(config)
    return server, worker, comm


def run_worker(model_path, config, task, data_path, local_rank, backend):
    server, worker, comm = build_job(config, task, data_path, model_path, backend)
    try:
        if server is not None:
            print_rank("Launching server")
            server.run()
        else:
            print_rank("Worker on node {}: process started".format(comm.rank))
            worker.run()
    finally:
        comm.close()
        import torch.distributed as dist
        if dist.is_initialized():
            try:
                dist.barrier()
                dist.destroy_process_group()
            except Exception:
                pass


def prepare_experiment(args):
    """Create ``<out>/<experiment>/{models,log}``, copy the yaml, init logging, load+validate the config."""
    exp = args.experiment or os.environ.get("FLUTE_EXPERIMENT") or \
        os.path.splitext(os.path.basename(args.config))[0]
    root = os.path.join(args.outputPath, exp)
    model_path, log_path = os.path.join(root, "models"), os.path.join(root, "log")
    os.makedirs(model_path, exist_ok=True)
    os.makedirs(log_path, exist_ok=True)
    if int(os.environ.get("RANK", 0)) == 0:
        shutil.copyfile(args.config, os.path.join(root, "FLUTE_config.yaml"))
    init_logging(log_path, loglevel=getattr(logging, str(args.loglevel).upper(), logging.INFO))
    if int(os.environ.get("RANK", 0)) == 0:
        get_run().attach_file(os.path.join(log_path, "metrics.jsonl"))
    with open(args.config) as f:
        config = FLUTEConfig.from_dict(yaml.safe_load(f))
    config["data_path"] = args.dataPath
    config["output_path"] = args.outputPath
    config["model_path"] = model_path
    config["experiment_name"] = exp
    config["client_config"]["task"] = args.task
    config["server_config"]["task"] = args.task
    config.validate()
    return config, model_path


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("-config")
    p.add_argument("-outputPath")
    p.add_argument("-dataPath", default=None)
    p.add_argument("-task", default=None, help="Define the task for the run")
    p.add_argument("-backend", default=None, help="Define the communication protocol")
    p.add_argument("-num_skip_decoding", default=-1, type=int, help="Skip decoding in unsupervised learning mode")
    p.add_argument("--local_rank", default=-1, type=int)
    p.add_argument("-experiment", default=None, help="experiment (output sub-directory) name")
    p.add_argument("-loglevel", default="INFO")
    args = p.parse_args(argv)
    if args.backend is None:
        args.backend = "nccl" if torch.cuda.is_available() else "gloo"
    assert args.backend in ["nccl", "gloo"], f"Backend {args.backend} not recognized, please select nccl or gloo"
    if args.dataPath is None:
        args.dataPath = get_run().input_datasets.get("input", "./")
    config, model_path = prepare_experiment(args)
    run_worker(model_path, config, args.task, args.dataPath, args.local_rank, args.backend)


if __name__ == "__main__":
    main()
