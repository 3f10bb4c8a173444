#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gpu8.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention"
DL_ATTN_X=0 run 60 python tools/bench_attn.py 16
run 60 python tools/bench_attn.py 16
run 600 python -m pytest tests/test_pipeline_gpu.py -q -x -k "tiny_pipeline or 512_4step"
run 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pool-e2e
