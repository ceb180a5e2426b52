#!/usr/bin/env bash
# end-of-round validation on ONE box with 2 GPUs: smoke(), N=1 bench (both arms), N=2 default transport
set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "exit $?" >> gpurun_out/final_smoke.log; tail -2 gpurun_out/final_smoke.log | cut -c1-200
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ours.log 2>&1; echo "exit $?" >> gpurun_out/final_bench_ours.log; grep '^{"metric"' gpurun_out/final_bench_ours.log | cut -c1-230
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/final_bench_ours_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/final_bench_ours_2gpu.log; grep '^{"metric"' gpurun_out/final_bench_ours_2gpu.log | cut -c1-230; grep -o '"parallelism[^,]*' gpurun_out/final_bench_ours_2gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/final_bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/final_bench_ref.log; grep '^{"impl"' gpurun_out/final_bench_ref.log | cut -c1-230
