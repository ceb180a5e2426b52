#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "exit $?" >> gpurun_out/bench_kernels.log
tail -30 gpurun_out/bench_kernels.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 40 -c 2 -o gpurun_out/prof_gemm python tools/bench_kernels.py --only gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_update_kernel -s 6 -c 1 -o gpurun_out/prof_update python tools/bench_kernels.py --only arena > gpurun_out/ncu_arena.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_dgrad_kernel -s 30 -c 1 -o gpurun_out/prof_dgrad python tools/bench_kernels.py --only conv > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep
