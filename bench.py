"""Headline benchmark: federated rounds/second for ResNet-18 on FedCIFAR-100 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N ...            # the unmodified reference from baseline/_ref

One "step" = one FL round of the benchmark config: 10 sampled clients × (100 samples, batch 20 ⇒ 5 local SGD steps,
clip 5.0, gradient statistics) + weighted aggregation + server SGD(lr 1.0) + model (re)distribution + the
per-round ``latest_model.tar`` checkpoint.  Synthetic Fed-CIFAR-100-shaped data (500 users × 100 × 32×32×3 uint8),
random-init weights.  For N > 1 launch under ``torch.distributed.run`` (one rank per GPU).

``value`` is measured through the public API (``OptimizationServer.begin_training / run_rounds``) with HBM-resident
shards.  The timed region is bracketed by ``Server.sync_nodes()`` (every rank drains its GPU, records a CUDA event,
barriers); each rank reports its own device time between its events and the maximum is used.  ``e2e`` repeats the K
rounds on all ranks with the engines in streaming mode: every round's client shards are copied host(pinned)→device
inside the timed region and the round's record table is read back (it always is); wall clock around ``run_rounds``.
``gpu_launches`` counts launches of this repo's own kernels in the timed region (rank 0).  Multi-GPU jobs run 10
set-up rounds before the W warm-up rounds (peer mappings, CUDA-graph capture on every rank).
``FLUTE_BENCH_CPU=1`` runs the same protocol on CPU/gloo with a tiny population (control-flow test, no number).
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
from bench_common import (BASELINE_PUBLISHED_ROUNDS_PER_SEC, HEADLINE_METRIC, ClockSampler, emit,  # noqa: E402
                          parse_args)


def _reference(args):
    """Re-exec into a clean interpreter so nothing of this repo is importable on the reference's path."""
    os.environ["PYTHONPATH"] = ""
    script = os.path.join(ROOT, "baseline", "run_reference.py")
    os.execv(sys.executable, [sys.executable, script] + sys.argv[1:])


class FlagshipJob:
    def __init__(self, server, worker, comm, config):
        self.server, self.worker, self.comm, self.config = server, worker, comm, config
        self._begun = False

    def run_round(self):
        if not self._begun:
            self.server.begin_training()
            self._begun = True
        return self.server.run_rounds(1)

    def close(self):
        if self._begun:
            self.server.end_training()


def make_config(args_task="cv_resnet_fedcifar100", n_clients_per_round=10, rounds=10 ** 6, norm="gn", comm="auto",
                resident=True, compute_dtype="fp32"):
    import yaml
    from msrflute_b200.core.config import FLUTEConfig
    with open(os.path.join(ROOT, "experiments", args_task, "config.yaml")) as f:
        raw = yaml.safe_load(f)
    sc = raw["server_config"]
    sc["max_iteration"] = rounds
    sc["num_clients_per_iteration"] = n_clients_per_round
    sc["val_freq"], sc["rec_freq"] = 10 ** 9, 10 ** 9      # rounds only (the reference arm does the same)
    sc["initial_val"], sc["initial_rec"] = False, False
    sc.setdefault("b200", {}).update({"comm": comm, "device_resident_data": resident,
                                      "wave_batched": os.environ.get("FLUTE_WAVE", "1") != "0"})
    raw["model_config"]["group_norm"] = 2 if norm == "gn" else 0
    raw["model_config"]["compute_dtype"] = compute_dtype
    return FLUTEConfig.from_dict(raw)


def build_flagship(n_clients_per_round=10, users=None, norm="gn", comm="auto", out_dir=None, resident=True,
                   compute_dtype="fp32"):
    import tempfile
    from msrflute_b200 import cli
    task = "cv_resnet_fedcifar100"
    config = make_config(task, n_clients_per_round, norm=norm, comm=comm, resident=resident,
                         compute_dtype=compute_dtype)
    out_dir = out_dir or tempfile.mkdtemp(prefix="flute_bench_")
    model_path = os.path.join(out_dir, "models")
    os.makedirs(model_path, exist_ok=True)
    config["data_path"], config["output_path"], config["model_path"] = out_dir, out_dir, model_path
    config["experiment_name"] = "bench"
    config["client_config"]["task"] = task
    config["server_config"]["task"] = task
    config.validate()
    if users is not None:                       # smaller synthetic population (smoke tests)
        from msrflute_b200.utils.dataloaders_utils import get_exp_dataset
        from msrflute_b200.data import synthetic
        ds_cls = get_exp_dataset(task)
        ds_cls.synthetic_train = staticmethod(lambda: synthetic.make_image_classification(users, 100, (32, 32, 3), 100, seed=5))
        ds_cls.synthetic_test = staticmethod(lambda: synthetic.make_image_classification(max(users // 5, 2), 100, (32, 32, 3), 100, seed=6))
    server, worker, comm_obj = cli.build_job(config, task, out_dir, model_path)
    return FlagshipJob(server, worker, comm_obj, config)


def main():
    args = parse_args()
    if args.impl == "reference":
        _reference(args)
        return
    import logging
    import time
    import torch
    os.chdir(ROOT)
    logging.getLogger().setLevel(logging.WARNING)
    cpu_dry_run = os.environ.get("FLUTE_BENCH_CPU") == "1"       # control-flow test of the multi-rank protocol (gloo)
    if not torch.cuda.is_available() and not cpu_dry_run:
        emit({"metric": HEADLINE_METRIC, "value": None, "unavailable": "no CUDA device"})
        return
    cuda = torch.cuda.is_available()
    from msrflute_b200.ops import _ext
    if cuda:
        _ext.load(required=True)
    from msrflute_b200.core.federated import Server
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    job = build_flagship(args.clients_per_round, users=(20 if cpu_dry_run else None), norm=args.norm, comm=args.comm,
                         compute_dtype=os.environ.get("FLUTE_BENCH_DTYPE", "fp32"))
    comm = job.comm
    eng = getattr(job.worker, "engine", None)
    do_e2e = not args.no_e2e and (eng is not None or cpu_dry_run)

    def region_ms(e0, e1, t0, t1):
        return e0.elapsed_time(e1) if cuda else (t1 - t0) * 1e3

    if rank != 0:
        # Workers live in the command loop; the server brackets the timed regions with sync_nodes() (every rank
        # drains its GPU, records a CUDA event, then barriers), so each rank knows its own device time per region.
        job.worker.run()
        w = job.worker
        ms = w.sync_region_ms(0, 1) if cuda else 0.0
        ms_e2e = w.sync_region_ms(2, 3) if (cuda and do_e2e) else 0.0
        comm.gather_objects({"ms": ms, "ms_e2e": ms_e2e, "h2d": int(getattr(w.engine, "h2d_bytes_last_round", 0) or 0)})
        comm.close()
        return

    server = job.server
    server.begin_training()
    if comm.size > 1:
        # setup, not warm-up: the first rounds of a multi-GPU job pay one-off costs (symmetric-memory peer mappings and
        # signal pads, NCCL channel set-up, CUDA-graph capture on every rank) that take ~10 rounds to disappear
        server.run_rounds(10)
    server.run_rounds(args.warmup)
    Server.sync_nodes()                                          # barrier + device synchronize on every rank
    sampler = ClockSampler(torch.cuda.current_device() if cuda else 0).start()
    n0 = _ext.LAUNCH_COUNTER["n"]
    if cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    loss = server.run_rounds(args.steps)
    if cuda:
        e1.record()
    Server.sync_nodes()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    launches = _ext.LAUNCH_COUNTER["n"] - n0
    dev_ms = region_ms(e0, e1, t0, t1) if cuda else (t1 - t0) * 1e3
    wall_ms = (t1 - t0) * 1e3

    # End-to-end variant: every engine streams its clients' shards host(pinned)->device inside the timed region and the
    # per-round record table is read back (it always is); timed by wall clock around the public run_rounds() call.
    e2e = None
    if do_e2e:
        Server.sync_nodes({"resident": False})                  # engines keep the pack in pinned host memory
        server.run_rounds(2)                                    # settle
        Server.sync_nodes()
        t2 = time.perf_counter()
        server.run_rounds(args.steps)
        Server.sync_nodes()
        e2e_s = time.perf_counter() - t2
        e2e = {"value": args.steps / e2e_s, "unit": "rounds/s",
               "h2d_bytes_per_step": int(getattr(eng, "h2d_bytes_last_round", 0) or 0),
               "d2h_bytes_per_step": int(getattr(eng, "d2h_bytes_last_round", 0) or 0),
               "note": "wall clock incl. host orchestration, async checkpoint snapshots, H2D of the sampled clients' "
                       "shards from pinned memory and D2H of the per-round loss/statistics table (rank 0's bytes; "
                       "all ranks stream their own shards)"}
    server.end_training()
    ms = max(dev_ms, 0.0)
    rank_ms = None
    if comm.size > 1:
        others = [o for o in comm.gather_objects({"ms": ms}) if isinstance(o, dict)]
        rank_ms = [round(float(o.get("ms", 0.0)), 3) for o in others]
        ms = max([ms] + [o.get("ms", 0.0) for o in others])
        if e2e is not None:
            e2e["h2d_bytes_per_step"] += sum(int(o.get("h2d", 0)) for o in others)
    value = args.steps / (ms / 1e3)
    emit({
        "metric": HEADLINE_METRIC, "value": value, "unit": "rounds/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": value / BASELINE_PUBLISHED_ROUNDS_PER_SEC,
        "dtype": "bf16" if os.environ.get("FLUTE_BENCH_DTYPE", "fp32") == "bf16" else "fp32 (tf32 tensor-core convs)",
        "data": "synthetic Fed-CIFAR-100 shape (500 users x 100 x 32x32x3 uint8), random-init weights",
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "impl": "ours",
        "last_train_loss": loss, "rank_ms": rank_ms if comm.size > 1 else None,
        "config": {"model": "ResNet-18 + {} (1000-way FC like the reference's RESNET), 11.7M params".format(
                       "GroupNorm(2 ch/group, per-group affine)" if args.norm == "gn" else "BatchNorm2d"),
                   "clients_per_round": args.clients_per_round, "client_batch": 20, "local_steps_per_client": 5,
                   "global_batch": args.clients_per_round * 100, "seq_len": None,
                   "parallelism": "fl-clients-over-{}gpu(s), {} transport".format(world, comm.kind),
                   "l2": "no explicit flush: each round streams {}x46.8 MB weight+grad arenas (>126 MB L2) and "
                         "re-samples clients".format(args.clients_per_round),
                   "checkpoint": "latest_model.tar snapshot every round (async writer, latest-wins, <=0.25 s stale)"},
    })
    comm.close()


if __name__ == "__main__":
    main()
