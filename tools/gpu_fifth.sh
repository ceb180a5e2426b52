#!/usr/bin/env bash
# tcgen05 conv path: correctness (both impls), micro-benchmark, flagship bench, one ncu capture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "slot_conv" -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -25 gpurun_out/pytest_conv.log
timeout 300 python tools/bench_kernels.py --only conv > gpurun_out/bench_conv.log 2>&1; echo "exit $?" >> gpurun_out/bench_conv.log
tail -12 gpurun_out/bench_conv.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours_tc.log 2>&1; echo "exit $?" >> gpurun_out/bench_ours_tc.log
tail -4 gpurun_out/bench_ours_tc.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 3 -o gpurun_out/prof_convtc python tools/bench_kernels.py --only conv > gpurun_out/ncu_convtc.log 2>&1
ls -la gpurun_out/*.ncu-rep
