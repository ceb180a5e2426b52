#!/usr/bin/env bash
# multi-GPU validation: N = $1 (default 2); every run is bounded (a hung rendezvous must not eat the GPU budget)
N=${1:-2}
set -x
mkdir -p gpurun_out
run() { timeout ${TMO:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 ${@:2}; }
run 29511 > gpurun_out/mg${N}_collective.log 2>&1; echo "exit $?" >> gpurun_out/mg${N}_collective.log; tail -2 gpurun_out/mg${N}_collective.log | cut -c1-900
FLUTE_COMM=symm run 29512 > gpurun_out/mg${N}_symm.log 2>&1; echo "exit $?" >> gpurun_out/mg${N}_symm.log; tail -2 gpurun_out/mg${N}_symm.log | cut -c1-900
