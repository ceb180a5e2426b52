#!/usr/bin/env bash
# strong-scaling record on one 8-GPU box: N = 8, 4 (and symm at 8); every run bounded
set -x
mkdir -p gpurun_out
run() { N=$1; PORT=$2; shift 2; timeout ${TMO:-170} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 30 --warmup 6 "$@"; }
run 8 29521 > gpurun_out/mg8_collective.log 2>&1; echo "exit $?" >> gpurun_out/mg8_collective.log; grep '^{"metric"' gpurun_out/mg8_collective.log | cut -c1-260
FLUTE_COMM=symm run 8 29522 > gpurun_out/mg8_symm.log 2>&1; echo "exit $?" >> gpurun_out/mg8_symm.log; grep '^{"metric"' gpurun_out/mg8_symm.log | cut -c1-260
CUDA_VISIBLE_DEVICES=0,1,2,3 run 4 29523 > gpurun_out/mg4_collective.log 2>&1; echo "exit $?" >> gpurun_out/mg4_collective.log; grep '^{"metric"' gpurun_out/mg4_collective.log | cut -c1-260
run 8 29524 --clients-per-round 80 > gpurun_out/mg8_weak80.log 2>&1; echo "exit $?" >> gpurun_out/mg8_weak80.log; grep '^{"metric"' gpurun_out/mg8_weak80.log | cut -c1-260
