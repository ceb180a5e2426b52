#!/usr/bin/env bash
# Round-2 evidence run (single GPU): bench line, per-launch table, ncu captures of the SlotNet GEMM kernel in its four
# epilogue forms, compute-sanitizer over the new kernels.  Everything lands in gpurun_out/; tools/summarise_profiles.py
# turns it into the tracked files under profiles/.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_1gpu.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 5 --sync-ckpt --no-e2e > gpurun_out/r2_bench_1gpu_sync_ckpt.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --norm bn --no-e2e > gpurun_out/r2_bench_1gpu_bn.log 2>&1
timeout 300 python tools/profile_slotnet.py 10 > gpurun_out/r2_profile_slotnet.log 2>&1
cp gpurun_out/slotnet_ops.txt gpurun_out/r2_slotnet_ops.txt
timeout 300 python tools/profile_round.py > gpurun_out/r2_profile_round.log 2>&1
cp gpurun_out/round_timeline.txt gpurun_out/r2_round_timeline.txt
NCU="ncu --set full --clock-control none --import-source on"
# ops: 4 fprop+GN (layer1), 15 fprop+GN (layer3), 26 dgrad+GN-bwd (fc), 39 wgrad (layer3), 2 stem fprop (E_STORE)
timeout 600 $NCU -k regex:sn_gemm_kernel -s 62 -c 5 -o gpurun_out/r2_prof_slotnet python tools/ncu_slotnet.py 4 15 26 39 2 > gpurun_out/r2_ncu_slotnet.log 2>&1
SANITIZE_TIMEOUT=900
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_slotnet_gemm_gpu.py tests/test_lstm_gpu.py tests/test_nn_kernels_gpu.py -x -q -p no:cacheprovider \
    -k "not vmap and not 80-256-256" > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_lstm_gpu.py tests/test_nn_kernels_gpu.py -x -q -p no:cacheprovider -k "matches_reference or batch_norm" \
    > gpurun_out/r2_sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2_sanitize_racecheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/r2_sanitize_*.log | tail -12
tail -1 gpurun_out/r2_bench_1gpu.log | cut -c1-300
