#!/usr/bin/env bash
# What CI runs (the reference's azure-pipelines.yml only scans credentials and never runs a test):
#   1. cross-compile every kernel for sm_100a and import the package,
#   2. the CPU test-suite (includes 2-rank gloo jobs and the four hello-world tasks),
#   3. on a GPU runner additionally: kernel numerics (-m gpu), the smoke round, a short benchmark.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/ -x -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests/ -x -q -m gpu
  python -c "import __graft_entry__ as g; g.smoke()"
  python bench.py --steps 10 --warmup 3
fi
