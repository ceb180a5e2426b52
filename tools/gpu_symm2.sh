#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 40 --warmup 12 --no-e2e; }
FLUTE_COMM=symm run 29531 > gpurun_out/mg2_symm.log 2>&1; echo "exit $?" >> gpurun_out/mg2_symm.log; grep '^{"metric"' gpurun_out/mg2_symm.log | cut -c1-250; grep -o '"rank_ms[^]]*]' gpurun_out/mg2_symm.log
run 29532 > gpurun_out/mg2_collective.log 2>&1; echo "exit $?" >> gpurun_out/mg2_collective.log; grep '^{"metric"' gpurun_out/mg2_collective.log | cut -c1-250; grep -o '"rank_ms[^]]*]' gpurun_out/mg2_collective.log
