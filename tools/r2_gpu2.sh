#!/bin/bash
# round 2, GPU call 2: stagger A/B of attention_pp + one ncu --set full capture of it
mkdir -p gpurun_out
L=gpurun_out/r2_gpu2.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
for st in 0 300 600 900 1200; do
  DL_ATTN_PP_STAGGER=$st run 60 python tools/bench_attn1.py
done
DL_ATTN_PP_POLY=0 DL_ATTN_PP_STAGGER=600 run 60 python tools/bench_attn1.py
DL_ATTN_PP_POLY=3 DL_ATTN_PP_STAGGER=600 run 60 python tools/bench_attn1.py
run 60 python tools/bench_attn1.py 8 9216 8 40
REPS=1 run 300 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 2 -c 1 -f -o gpurun_out/r2_attn_pp python tools/bench_attn1.py
