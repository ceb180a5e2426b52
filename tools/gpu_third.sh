#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_slotconv.log 2>&1; echo "exit $?" >> gpurun_out/bench_ours_slotconv.log
tail -4 gpurun_out/bench_ours_slotconv.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --gpus 1 --steps 2 --warmup 4 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launches.log
tail -2 gpurun_out/ncu_launches.log
