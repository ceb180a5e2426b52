#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for t in mlm_bert nlg_gru; do
  timeout 300 python e2e_trainer.py -dataPath /tmp/hw_$t -outputPath /tmp/hw_$t -config testing/hello_world_$t.yaml -task $t -experiment hello > gpurun_out/hello_gpu_$t.log 2>&1; echo "exit $?" >> gpurun_out/hello_gpu_$t.log
  grep -E "tcgen05|parameters|Training loss|exit|Error|error" gpurun_out/hello_gpu_$t.log | tail -6 | cut -c1-200
done
