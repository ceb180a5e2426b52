#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "slot_conv or engine_round or slot_batched" -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -12 gpurun_out/pytest_conv.log
FLUTE_CONV_IMPL=auto timeout 600 python bench.py --steps 30 --warmup 6 > gpurun_out/b_auto.log 2>&1; tail -1 gpurun_out/b_auto.log | cut -c1-220
ROUNDS=6 timeout 600 python tools/profile_round.py > gpurun_out/profile_round.log 2>&1; head -12 gpurun_out/round_timeline.txt | cut -c1-160
timeout 600 python tools/profile_host.py > gpurun_out/profile_host.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
