#!/bin/bash
# round 2, GPU call 1: parity + A/B of the two-query-tile attention kernel (attention_pp.cu),
# the mode-6 variant left from round 1, then a short bench.
mkdir -p gpurun_out
L=gpurun_out/r2_gpu1.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $L 2>&1
run 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 
DL_ATTN_PP=0 run 60 python tools/bench_attn.py 16
DL_ATTN_PP_POLY=0 run 60 python tools/bench_attn.py 16
DL_ATTN_PP_POLY=2 run 60 python tools/bench_attn.py 16
DL_ATTN_PP_POLY=3 run 60 python tools/bench_attn.py 16
DL_ATTN_PP_POLY=4 run 60 python tools/bench_attn.py 16
DL_ATTN_MODE=6 run 120 python -m pytest tests/test_kernels_gpu.py -q -x -k attention_tc
DL_ATTN_MODE=6 run 60 python tools/bench_attn.py 16
DL_ATTN_MODE=6 DL_ATTN_POLY=0 run 60 python tools/bench_attn.py 16
run 300 python -m pytest tests/test_pipeline_gpu.py -q -x -k "tiny_pipeline or 512_4step"
run 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
