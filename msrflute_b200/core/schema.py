"""Rule schema for the six-section FL YAML configuration.

Same key names / required-ness / defaults as the reference's
``core/schema.py`` (top-level sections at ``schema.py:10,38,55,71,76,223``;
defaults ``schema.py:106-111,123-134``) so every reference YAML validates
unchanged, but generated programmatically: the three ``data_config`` blocks and
the four optimizer blocks share builders instead of being spelled out.

New, B200-specific keys live under ``server_config.b200`` (all optional, see
``B200_KEYS``) and are ignored by the reference.
"""
from __future__ import annotations

OPTIMIZER_TYPES = ["sgd", "adam", "adamax", "lars", "LarsSGD", "lamb", "adamW"]
SERVER_TYPES = ["model_optimization", "personalization"]
CLIENT_TYPES = ["optimization", "gradient_computation"]


def _opt(t, **kw):
    return dict(required=False, type=t, **kw)


def _req(t, **kw):
    return dict(required=True, type=t, **kw)


def _section(schema, required=True, allow_unknown=True, **kw):
    return dict(required=required, type="dict", allow_unknown=allow_unknown, schema=schema, **kw)


def _loader_common():
    """Keys shared by every train/val/test data block."""
    return {
        "batch_size": _opt("integer", default=40),
        "tokenizer_type": _opt("string"),
        "prepend_datapath": _opt("boolean", default=False),
        "vocab_dict": _opt("string", nullable=True),
        "pin_memory": _opt("boolean", default=True),
        "num_workers": _opt("integer", default=1),
        "num_frames": _opt("integer", default=0),
        "max_batch_size": _opt("integer", default=0),
        "max_num_words": _opt("integer"),
        "max_grad_norm": _opt("float", default=5.0),
        "unsorted_batch": _opt("boolean", default=False),
    }


def _data_block(data_key=None, data_required=True, extra=None, required=True):
    block = _loader_common()
    if data_key is not None:
        block[data_key] = dict(required=data_required, type="string", nullable=True)
    block.update(extra or {})
    return _section(block, required=required)


def _optimizer_block(lr_required=True, required=True, extra=None):
    block = {
        "type": _req("string", allowed=list(OPTIMIZER_TYPES)),
        "lr": dict(required=lr_required, type="float"),
        "weight_decay": _opt("float"),
    }
    block.update(extra or {})
    return _section(block, required=required)


def _annealing_block(strict):
    return _section({
        "type": _req("string"),
        "step_interval": _req("string"),
        "gamma": dict(required=strict, type="float"),
        "step_size": dict(required=strict, type="integer"),
    }, required=strict)


B200_KEYS = {
    # transport for the weight broadcast / pseudo-gradient gather:
    #   auto  -> symm on multi-GPU NCCL jobs, collective otherwise
    #   symm  -> hand-written P2P/NVLS kernels over symmetric memory
    #   collective -> torch.distributed reduce/broadcast (gloo or nccl baseline)
    #   p2p   -> reference-style per-client point-to-point delivery
    "comm": _opt("string", allowed=["auto", "symm", "collective", "p2p"], default="auto"),
    "cuda_graphs": _opt("boolean", default=True),
    "max_concurrent_clients": _opt("integer", default=16),
    "device_resident_data": _opt("boolean", default=True),
    "compute_dtype": _opt("string", allowed=["fp32", "bf16"], default="bf16"),
    "server_is_worker": _opt("boolean", default=True),
    "dispatch": _opt("string", allowed=["static_lpt", "round_robin", "dynamic", "work_queue"], default="static_lpt"),
    "seed": _opt("integer", default=0),
    "async_checkpoint": _opt("boolean"),
    "device_engine": _opt("boolean", default=True),
    "wave_batched": _opt("boolean", default=True),
}


def build_schema():
    cache = {"cache_dir": _opt("string")}
    server_data = _section({
        "val": _data_block("val_data", extra=cache),
        "test": _data_block("test_data", extra=cache),
        "train": _data_block(None, required=False, extra={
            "train_data_server": _opt("string"),
            "desired_max_samples": _opt("integer"),
            **cache,
        }),
    }, keysrules={"forbidden": ["num_clients"]})

    client_data = _section({
        "train": _data_block("list_of_train_data"),
    }, keysrules={"forbidden": ["num_clients"]})

    bert = _section({
        "loader_type": _opt("string"),
        "model": _section({
            "model_name_or_path": _opt("string"),
            "model_name": _req("string"),
            "process_line_by_line": _req("boolean"),
        }),
    }, required=False)

    return {
        "model_config": _section({
            "model_type": _req("string"),
            "model_folder": _req("string"),
            "BERT": bert,
        }),
        "dp_config": _section({
            "enable_local_dp": _req("boolean"),
            "enable_global_dp": _opt("boolean"),
            "eps": _opt("float"),
            "delta": _opt("float"),
            "global_sigma": _opt("float"),
            "max_grad": _opt("float"),
            "max_weight": _opt("float"),
            "weight_scaler": _opt("float"),
            "min_weight": _opt("float"),
        }),
        "privacy_metrics_config": _section({
            "apply_metrics": _req("boolean"),
            "apply_indices_extraction": _opt("boolean"),
            "allowed_word_rank": _opt("integer"),
            "apply_leakage_metric": _opt("boolean"),
            "max_leakage": _opt("float"),
            "adaptive_leakage_threshold": _opt("float"),
            "is_leakage_weighted": _opt("boolean"),
            "attacker_optimizer_config": dict(required=False, type="dict", allow_unknown=True),
        }),
        "strategy": _req("string"),
        "server_config": _section({
            "wantRL": _req("boolean"),
            "RL": _opt("dict"),
            "resume_from_checkpoint": _req("boolean"),
            "do_profiling": _req("boolean"),
            "optimizer_config": _optimizer_block(),
            "annealing_config": _annealing_block(strict=True),
            "val_freq": _opt("integer", default=1),
            "rec_freq": _opt("integer", default=8),
            "initial_val": _opt("boolean", default=True),
            "initial_rec": _opt("boolean", default=False),
            "max_iteration": _opt("integer", default=10000),
            # the reference accepts "lo,hi" strings at run time (server.py:84-86)
            "num_clients_per_iteration": dict(required=False, type=["integer", "string"], default=1),
            "data_config": server_data,
            "type": _opt("string", allowed=list(SERVER_TYPES), default="model_optimization"),
            "aggregate_median": _opt("string"),
            "initial_lr_client": _req("float"),
            "lr_decay_factor": _req("float"),
            "weight_train_loss": _req("string"),
            "best_model_criterion": _opt("string", default="loss"),
            "fall_back_to_best_model": _opt("boolean", default=False),
            "softmax_beta": _req("float"),
            "server_replay_config": dict(required=False, type="dict", allow_unknown=True, schema={
                "server_iterations": _req("integer"),
                "optimizer_config": _optimizer_block(extra={"amsgrad": _opt("boolean")}),
            }),
            "nbest_task_scheduler": dict(required=False, type="dict", schema={
                "num_tasks": dict(required=True, type=["integer", "list"]),
                "iteration_per_task": dict(required=True, type=["integer", "list"]),
            }),
            "b200": _section(dict(B200_KEYS), required=False),
        }),
        "client_config": _section({
            "meta_learning": _opt("string"),
            "stats_on_smooth_grad": _opt("boolean"),
            "ignore_subtask": _req("boolean"),
            "num_skips_threshold": _opt("integer"),
            "copying_train_data": _opt("boolean"),
            "do_profiling": _req("boolean"),
            "data_config": client_data,
            "type": _opt("string", allowed=list(CLIENT_TYPES), default="gradient_computation"),
            "meta_optimizer_config": _section({
                "type": _req("string", allowed=list(OPTIMIZER_TYPES)),
                "lr": _req("float"),
            }, required=False),
            "optimizer_config": _optimizer_block(lr_required=False),
            "annealing_config": _annealing_block(strict=False),
            "ss_config": dict(required=False, type="dict", allow_unknown=True, nullable=True),
            # B200: capture the generic client mini-batch step once into a CUDA graph (core/graphed.py)
            "graphed_step": _opt("boolean"),
        }),
    }


SCHEMA = build_schema()
