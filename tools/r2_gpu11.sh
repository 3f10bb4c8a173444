#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gpu11.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 2400 python -m pytest tests/ -q -x -m gpu
DL_VAE_CONV_OUT_TAPSUM=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
