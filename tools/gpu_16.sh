#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "slot_conv or slot_batched or engine_round" > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -6 gpurun_out/pytest_conv.log
timeout 400 python bench.py --steps 30 --warmup 6 > gpurun_out/b_auto.log 2>&1; tail -1 gpurun_out/b_auto.log | cut -c1-220
