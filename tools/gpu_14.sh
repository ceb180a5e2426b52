#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -6 gpurun_out/pytest_gemm.log
timeout 300 python tools/bench_kernels.py --only gemm > gpurun_out/bench_gemm.log 2>&1; echo "exit $?" >> gpurun_out/bench_gemm.log
tail -9 gpurun_out/bench_gemm.log | cut -c1-420
