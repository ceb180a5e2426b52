set -x
mkdir -p gpurun_out
FLUTE_MG_FULL=1 timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_multigpu_tests_8gpu.log
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 100 --warmup 5 --no-e2e 2>/dev/null | grep "^{" | tail -1 > gpurun_out/r2_bench_${N}gpu.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e --clients-per-round 80 2>/dev/null | grep "^{" | tail -1 > gpurun_out/r2_bench_8gpu_weak80.json
cat gpurun_out/r2_multigpu_tests_8gpu.log
for f in gpurun_out/r2_bench_8gpu.json gpurun_out/r2_bench_4gpu.json gpurun_out/r2_bench_2gpu.json gpurun_out/r2_bench_8gpu_weak80.json; do cut -c1-420 $f; echo; done
