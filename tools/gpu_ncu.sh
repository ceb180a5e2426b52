#!/usr/bin/env bash
# one `ncu --set full` capture per hot kernel (single GPU, never under a timed bench), reports land in gpurun_out/
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv_tc_kernel -s 200 -c 3 -o gpurun_out/prof_convtc2 python bench.py --steps 2 --warmup 4 --no-e2e > gpurun_out/ncu_convtc2.log 2>&1
timeout 300 $NCU -k regex:gemm_bf16_tn_persistent -s 30 -c 1 -o gpurun_out/prof_gemm_persistent python tools/bench_kernels.py --only gemm > gpurun_out/ncu_gemm2.log 2>&1
ls -la gpurun_out/*.ncu-rep
