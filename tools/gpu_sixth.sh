#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
FLUTE_CONV_IMPL=fma timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/b_fma.log 2>&1; tail -1 gpurun_out/b_fma.log | cut -c1-200
FLUTE_CONV_IMPL=tcgen05 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/b_tc.log 2>&1; tail -1 gpurun_out/b_tc.log | cut -c1-200
FLUTE_CONV_IMPL=auto timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/b_auto.log 2>&1; tail -1 gpurun_out/b_auto.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1200 --csv --log-file gpurun_out/launches_tc.csv python bench.py --gpus 1 --steps 2 --warmup 4 --no-e2e > gpurun_out/ncu_launches_tc.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launches_tc.log
tail -2 gpurun_out/ncu_launches_tc.log | cut -c1-300
