#!/usr/bin/env bash
# Race / memory checking of the hand-written kernels (the reference has none).
#   memcheck  : out-of-bounds / misaligned accesses (arena tails, split-K tiles, segment tables)
#   racecheck : shared-memory hazards inside a CTA (block reductions, staged tiles)
#   synccheck : barrier misuse
# Usage (GPU box):  bash tools/sanitize.sh [memcheck|racecheck|synccheck] [pytest -k expression]
TOOL=${1:-memcheck}
EXPR=${2:-"arena or group_norm or misc or quantize or softmax_ce or cosine or local_dp"}
mkdir -p gpurun_out
timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 60 \
    python -m pytest tests/test_gpu_kernels.py -x -q -k "$EXPR" -p no:cacheprovider \
    > gpurun_out/sanitize_${TOOL}.log 2>&1
echo "compute-sanitizer $TOOL exit $?" | tee -a gpurun_out/sanitize_${TOOL}.log
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${TOOL}.log | tail -5
