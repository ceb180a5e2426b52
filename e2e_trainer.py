"""FLUTE-compatible launcher.

    python -m torch.distributed.run --nproc_per_node=N e2e_trainer.py \
        -dataPath <data> -outputPath <out> -config <yaml> -task <task> -backend nccl|gloo

Same flags and output tree as the reference's ``e2e_trainer.py``; the implementation is in
``msrflute_b200/cli.py``.
"""
from msrflute_b200.cli import main, run_worker, log_run_properties  # noqa: F401

if __name__ == "__main__":
    main()
