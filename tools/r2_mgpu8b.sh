#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_mgpu8b.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nproc >> $L
B200_PNG=gpu run 600 python bench.py --pool-workers 8 --steps 12
B200_PNG=gpu B200_PNG_THREADS=32 run 600 python bench.py --pool-workers 8 --steps 12
