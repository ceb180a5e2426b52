#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "slot_conv or slot_batched" > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -4 gpurun_out/pytest_conv.log
timeout 600 python bench.py --steps 30 --warmup 6 > gpurun_out/b_auto.log 2>&1; tail -1 gpurun_out/b_auto.log | cut -c1-220
ROUNDS=6 timeout 600 python tools/profile_round.py > gpurun_out/profile_round.log 2>&1; head -2 gpurun_out/round_timeline.txt; grep -A 12 "top kernels" gpurun_out/round_timeline.txt | cut -c1-120
