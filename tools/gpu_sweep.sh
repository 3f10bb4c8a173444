#!/bin/bash
# Run each GPU test function in its own process (a device fault in one must not mask the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
for t in "$@"; do
  echo "=== $t"
  timeout 600 python -m pytest -q -x "$t" -m gpu -p no:cacheprovider 2>&1 | tail -25
done
