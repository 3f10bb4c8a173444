#!/bin/bash
# First GPU run of the EXPERIMENTAL three-CTAs-per-SM attention variant (DL_ATTN_MODE=6: 64-key tile,
# one softmax thread per row).  Parity first, then the micro-benchmark against the default.
mkdir -p gpurun_out
L=gpurun_out/attn_mode6.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
DL_ATTN_MODE=6 run 60 python -m pytest tests/test_kernels_gpu.py -q -x -k attention_tc
DL_ATTN_MODE=6 run 30 python tools/bench_attn.py 16
DL_ATTN_MODE=6 DL_ATTN_POLY=0 run 30 python tools/bench_attn.py 16
run 30 python tools/bench_attn.py 16
DL_ATTN_MODE=6 run 60 python -m pytest tests/test_pipeline_gpu.py -q -s -k "tiny_pipeline or 512_4step"
