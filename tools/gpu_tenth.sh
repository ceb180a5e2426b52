#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "quantize or local_dp or softmax_ce or cosine" > gpurun_out/pytest_misc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_misc.log
tail -12 gpurun_out/pytest_misc.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
