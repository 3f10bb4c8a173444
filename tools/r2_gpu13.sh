#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gpu13.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_wide"
run 300 python tools/bench_vae_attn.py 16 512
run 300 python tools/bench_vae_attn.py 8 768
run 300 python tools/bench_vae_attn.py 1 1024
