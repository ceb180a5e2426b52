#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "compact" > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -4 gpurun_out/pytest_conv.log
FLUTE_CONV_TC_MIN_ROWS=16 timeout 400 python bench.py --steps 30 --warmup 6 --no-e2e > gpurun_out/b_tc16.log 2>&1; tail -1 gpurun_out/b_tc16.log | cut -c1-200
timeout 400 python bench.py --steps 30 --warmup 6 --no-e2e > gpurun_out/b_tc48.log 2>&1; tail -1 gpurun_out/b_tc48.log | cut -c1-200
