#!/usr/bin/env bash
# first GPU session: tests, short bench (ours + reference), launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 5 > gpurun_out/bench_ours.log 2>&1; echo "exit $?" >> gpurun_out/bench_ours.log
tail -5 gpurun_out/bench_ours.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref.log
tail -3 gpurun_out/bench_ref.log
