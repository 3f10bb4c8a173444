#!/bin/bash
# PDL (programmatic dependent launch) A/B: full GPU suite with PDL on, then bench with DL_PDL=0 / 1 in one call
mkdir -p gpurun_out
L=gpurun_out/r2_gpu12.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 2400 python -m pytest tests/ -q -x -m gpu
DL_PDL=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
DL_PDL=1 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
DL_PDL=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
DL_PDL=1 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
