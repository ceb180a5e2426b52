#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "slot_conv or engine_round or slot_batched" -q > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_conv.log
tail -12 gpurun_out/pytest_conv.log
FLUTE_CONV_IMPL=auto timeout 600 python bench.py --steps 30 --warmup 6 > gpurun_out/b_auto.log 2>&1; tail -1 gpurun_out/b_auto.log | cut -c1-220
ROUNDS=6 timeout 600 python tools/profile_round.py > gpurun_out/profile_round.log 2>&1; head -4 gpurun_out/round_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches_tc.csv python bench.py --gpus 1 --steps 2 --warmup 4 --no-e2e > gpurun_out/ncu_launches_tc.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launches_tc.log
