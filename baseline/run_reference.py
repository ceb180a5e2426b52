"""Reference arm of the benchmark: runs the UNMODIFIED microsoft/msrflute from ``baseline/_ref`` through its own
entry point (``e2e_trainer.py`` ``__main__``: argparse → FLUTEConfig → ``run_worker`` → OptimizationServer /
Worker) on synthetic Fed-CIFAR-100-shaped data, and times its FL rounds.

Nothing of msrflute_b200 is on this path.  What is added around the reference:
* ``baseline/shims``: offline stand-ins for the uninstallable imports (azureml, cerberus, easydict, h5py, wget);
* the data files its own reader opens (``./data/fed_cifar100/*.h5``) pre-seeded with synthetic arrays;
* a callback on the azureml shim that timestamps the reference's own per-round ``secsPerRoundTotal`` log call
  (``core/server.py:499``) after a device synchronize — that is how rounds are delimited without touching its code.
"""
import json
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, HERE)
from bench_common import (BASELINE_PUBLISHED, BASELINE_PUBLISHED_ROUNDS_PER_SEC, HEADLINE_METRIC, TASKS, ClockSampler, emit,  # noqa: E402
                          parse_args)


def seed_data(root):
    import numpy as np
    d = os.path.join(root, "data", "fed_cifar100")
    os.makedirs(d, exist_ok=True)
    tar = os.path.join(root, "data", "fed_cifar100.tar.bz2")
    if not os.path.exists(tar):
        open(tar, "wb").close()                       # its presence makes download_files() skip wget
    for fname, users, seed in (("fed_cifar100_train.h5", 500, 5), ("fed_cifar100_test.h5", 100, 6)):
        path = os.path.join(d, fname)
        if os.path.exists(path):
            continue
        rng = np.random.default_rng(seed)
        protos = rng.random((100, 32, 32, 3)).astype(np.float32)
        n = 100
        labels = rng.integers(0, 100, size=(users, n))
        imgs = np.empty((users * n, 32, 32, 3), dtype=np.uint8)
        for u in range(users):
            x = 0.5 * protos[labels[u]] + 0.5 * rng.random((n, 32, 32, 3), dtype=np.float32)
            imgs[u * n:(u + 1) * n] = (x * 255).astype(np.uint8)
        with open(path, "wb") as f:
            np.savez(f, users=np.array(["u{:05d}".format(u) for u in range(users)]),
                     offsets=np.arange(users + 1) * n, f_image=imgs, f_label=labels.reshape(-1).astype(np.int64))


def seed_femnist(root):
    """FedEMNIST-shaped files for the reference's own reader (``./data/femnist/fed_emnist_*.h5`` through the h5py shim):
    3400 train users x ~100 samples, 28x28 float pixels, 62 classes — the shape ``experiments/cv_cnn_femnist`` of this
    repo synthesises for its own arm."""
    import numpy as np
    d = os.path.join(root, "data", "femnist")
    os.makedirs(d, exist_ok=True)
    tar = os.path.join(root, "data", "fed_emnist.tar.bz2")
    if not os.path.exists(tar):
        open(tar, "wb").close()
    for fname, users, mean, seed in (("fed_emnist_train.h5", 3400, 100, 3), ("fed_emnist_test.h5", 340, 40, 4)):
        path = os.path.join(d, fname)
        if os.path.exists(path):
            continue
        rng = np.random.default_rng(seed)
        protos = np.random.default_rng(1234).random((62, 28, 28)).astype(np.float32)
        sizes = np.maximum(2, rng.poisson(mean, size=users))
        off = np.concatenate([[0], np.cumsum(sizes)])
        labels = rng.integers(0, 62, size=off[-1])
        px = (0.5 * protos[labels] + 0.5 * rng.random((off[-1], 28, 28), dtype=np.float32)).astype(np.float32)
        with open(path, "wb") as f:
            np.savez(f, users=np.array(["f{:05d}".format(u) for u in range(users)]), offsets=off, f_pixels=px,
                     f_label=labels.astype(np.int64))


def seed_shakespeare(root):
    """Shakespeare-shaped files: 715 train users whose ``snippets`` tokenise (by the reference's own ``preprocess``)
    into ~50 sequences of 80 characters each."""
    import numpy as np
    d = os.path.join(root, "data", "fed_shakespeare")
    os.makedirs(d, exist_ok=True)
    tar = os.path.join(root, "data", "shakespeare.tar.bz2")
    if not os.path.exists(tar):
        open(tar, "wb").close()
    vocab = list("dhlptx@DHLPTX $(,048cgkoswCGKOSW[_#'/37;?bfjnrvzBFJNRVZ&*.26:aeimquyAEIMQUY]!%)-159")
    for fname, users, mean, seed in (("shakespeare_train.h5", 715, 50, 7), ("shakespeare_test.h5", 100, 30, 8)):
        path = os.path.join(d, fname)
        if os.path.exists(path):
            continue
        rng = np.random.default_rng(seed)
        sizes = np.maximum(2, rng.poisson(mean, size=users))
        snippets = []
        for n in sizes:
            chars = rng.integers(0, len(vocab), size=int(n) * 81 - 2)       # + <bos>, <eos> = n full sequences
            snippets.append("".join(vocab[c] for c in chars).encode("utf8"))
        arr = np.empty(len(snippets), dtype=object)
        arr[:] = snippets
        with open(path, "wb") as f:
            np.savez(f, users=np.array(["s{:05d}".format(u) for u in range(users)]), offsets=np.arange(users + 1),
                     f_snippets=arr)


SEEDERS = {"cv_resnet_fedcifar100": ("fed_cifar100", "fed_cifar100_test.h5"),
           "cv_cnn_femnist": ("femnist", "fed_emnist_test.h5"),
           "nlp_rnn_fedshakespeare": ("fed_shakespeare", "shakespeare_test.h5")}


def main():
    args = parse_args()
    if args.clients_per_round is None:
        args.clients_per_round = TASKS[args.task]["clients_per_round"]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    os.environ.setdefault("RANK", str(rank))
    os.environ.setdefault("WORLD_SIZE", str(world))
    os.environ.setdefault("LOCAL_RANK", str(rank))
    if not os.path.isdir(os.path.join(REF, "core")):
        if rank == 0:
            emit({"impl": "reference", "unavailable": "baseline/_ref missing: run baseline/install_reference.sh"})
        return
    os.chdir(REF)                                      # the reference resolves ./core/schema.py and ./experiments/…
    sys.path[:0] = [os.path.join(HERE, "shims"), REF]
    import torch
    backend = os.environ.get("FLUTE_REF_BACKEND", "nccl")     # "gloo" = CPU smoke test of the shims only
    if not torch.cuda.is_available() and backend != "gloo":
        if rank == 0:
            emit({"impl": "reference", "unavailable": "no CUDA device: the reference's NCCL path needs GPUs"})
        return
    if args.task not in SEEDERS:
        if rank == 0:
            emit({"impl": "reference", "metric": TASKS[args.task]["metric"],
                  "unavailable": "no offline data seeder for the reference's {} reader (needs HF model/tokenizer "
                                 "downloads)".format(args.task)})
        return
    if rank == 0:
        {"cv_resnet_fedcifar100": seed_data, "cv_cnn_femnist": seed_femnist,
         "nlp_rnn_fedshakespeare": seed_shakespeare}[args.task](REF)
    else:
        sub, last = SEEDERS[args.task]
        while not os.path.exists(os.path.join(REF, "data", sub, last)):
            time.sleep(0.5)
        time.sleep(1.0)

    import yaml
    with open(os.path.join(REF, "experiments", args.task, "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    total = args.warmup + args.steps
    sc = cfg["server_config"]
    sc["max_iteration"] = total
    sc["num_clients_per_iteration"] = args.clients_per_round
    sc["val_freq"], sc["rec_freq"] = 10 ** 9, 10 ** 9     # rounds only, like our arm: no eval inside the timed region
    sc["initial_val"], sc["initial_rec"] = False, False
    for section, upd in (TASKS[args.task].get("overrides") or {}).items():     # same algorithmic config as our arm
        tgt, keys = cfg, section.split(".")
        for k in keys[:-1]:
            tgt = tgt.setdefault(k, {})
        tgt[keys[-1]] = upd
    out_dir = os.path.join("/tmp", "flute_ref_bench_{}".format(os.environ.get("MASTER_PORT", "0")))
    os.makedirs(out_dir, exist_ok=True)
    cfg_path = os.path.join(out_dir, "bench_{}.yaml".format(args.task))
    if rank == 0:
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
    else:
        while not os.path.exists(cfg_path):
            time.sleep(0.2)
        time.sleep(0.5)

    from azureml.core import Run
    marks = []
    state = {"sampler": None, "clocks": None}

    def on_log(key, value):
        if key != "secsPerRoundTotal":
            return
        ev = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            ev.synchronize()
        marks.append((time.perf_counter(), ev))
        if len(marks) == args.warmup:
            state["sampler"] = ClockSampler(int(os.environ.get("LOCAL_RANK", 0))).start()
        if len(marks) == total and state["sampler"] is not None:
            state["clocks"] = state["sampler"].stop()

    if rank == 0:
        Run.get_context().callbacks.append(on_log)
    sys.argv = ["e2e_trainer.py", "-dataPath", os.path.join(REF, "data"), "-outputPath", out_dir, "-config", cfg_path,
                "-task", args.task, "-backend", backend]
    t0 = time.perf_counter()
    runpy.run_path(os.path.join(REF, "e2e_trainer.py"), run_name="__main__")
    if rank != 0:
        return
    if len(marks) < total or args.warmup < 1:
        emit({"impl": "reference", "unavailable": "reference finished {} of {} rounds".format(len(marks), total)})
        return
    (w0, e0), (w1, e1) = marks[args.warmup - 1], marks[total - 1]
    wall_ms = (w1 - w0) * 1e3
    ms = e0.elapsed_time(e1) if e0 is not None else wall_ms      # CUDA events on rank 0 (it waits for every worker)
    value = args.steps / (ms / 1e3)
    spec = TASKS[args.task]
    flagship = args.task == "cv_resnet_fedcifar100"
    if not flagship:
        emit({"impl": "reference", "metric": spec["metric"], "value": value, "unit": "rounds/s", "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
              "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": (value / BASELINE_PUBLISHED[args.task]) if BASELINE_PUBLISHED.get(args.task) else None,
              "dtype": "fp32", "data": spec["data"], "clocks": state["clocks"],
              "e2e": {"value": args.steps / (wall_ms / 1e3), "unit": "rounds/s", "h2d_bytes_per_step": None,
                      "d2h_bytes_per_step": None,
                      "note": "wall clock between the reference's own per-round log calls (it moves every mini-batch "
                              "host->device and every gradient device->host itself)"},
              "gpu_launches": None,
              "config": dict(spec["config"], clients_per_round=args.clients_per_round,
                             parallelism="server+{}workers".format(max(world - 1, 1)) if world > 1
                             else "single-gpu thread path",
                             note="reference config.yaml as shipped (strategy/quantization overrides of our arm "
                                  "are applied where the reference supports them)")})
        return
    emit({"impl": "reference", "metric": HEADLINE_METRIC, "value": value, "unit": "rounds/s", "n_gpus": world,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
          "higher_is_better": True, "scaling": "strong", "vs_baseline": value / BASELINE_PUBLISHED_ROUNDS_PER_SEC,
          "dtype": "fp32(tf32 conv)", "data": "synthetic Fed-CIFAR-100 shape (500 users x 100 x 32x32x3 uint8), random-init weights",
          "clocks": state["clocks"],
          # the reference IS end to end: every mini-batch is moved host->device inside model.loss() and every step
          # copies the loss (and all gradients, trainer.py:277-288) back; so its e2e number is its wall clock
          "e2e": {"value": args.steps / (wall_ms / 1e3), "unit": "rounds/s",
                  "h2d_bytes_per_step": args.clients_per_round * 100 * (3 * 32 * 32 * 4 + 8),
                  "d2h_bytes_per_step": args.clients_per_round * (5 * (11689512 * 4 + 4) + 11689512 * 4),
                  "note": "wall clock between the reference's own per-round log calls; bytes counted from its code path "
                          "(float32 batches H2D per step; per-step gradient statistics + loss and the per-client "
                          "payload D2H)"},
          "gpu_launches": None,
          "config": {"model": "reference RESNET as shipped (resnet18, BatchNorm2d, 1000-way FC)",
                     "clients_per_round": args.clients_per_round, "client_batch": 20, "local_steps_per_client": 5,
                     "global_batch": args.clients_per_round * 100, "seq_len": None,
                     "parallelism": "server+{}workers".format(max(world - 1, 1)) if world > 1 else "single-gpu thread path",
                     "l2": "inputs re-sampled every round; per-round working set (10 model replicas' traffic) > L2",
                     "total_s_incl_startup": time.perf_counter() - t0}})


if __name__ == "__main__":
    main()
