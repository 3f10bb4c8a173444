#!/bin/bash
# A/B of the polynomial-exp2 attention variants (DL_ATTN_POLY) — micro-benchmark, kernel parity,
# pipeline parity and the whole step.  Writes gpurun_out/attn_poly.log.
mkdir -p gpurun_out
L=gpurun_out/attn_poly.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
DL_ATTN_POLY=2 run 25 python tools/bench_attn.py 16
DL_ATTN_POLY=0 run 25 python tools/bench_attn.py 16
DL_ATTN_POLY=2 run 40 python -m pytest tests/test_kernels_gpu.py -q -k attention_tc
DL_ATTN_POLY=2 run 40 python bench.py --no-cpu-baseline --steps 5
DL_ATTN_POLY=3 run 25 python tools/bench_attn.py 16
DL_ATTN_POLY=2 run 40 python -m pytest tests/test_pipeline_gpu.py -q -s -k "tiny_pipeline or 512_4step"
