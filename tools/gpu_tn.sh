#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "slot_conv or slot_batched or engine_round" > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -5 gpurun_out/pytest_conv.log
timeout 300 python bench.py --steps 30 --warmup 6 --no-e2e > gpurun_out/b_tnauto.log 2>&1; tail -1 gpurun_out/b_tnauto.log | cut -c1-200
FLUTE_CONV_TN=64 timeout 300 python bench.py --steps 30 --warmup 6 --no-e2e > gpurun_out/b_tn64.log 2>&1; tail -1 gpurun_out/b_tn64.log | cut -c1-200
FLUTE_CONV_TN=128 timeout 300 python bench.py --steps 30 --warmup 6 --no-e2e > gpurun_out/b_tn128.log 2>&1; tail -1 gpurun_out/b_tn128.log | cut -c1-200
