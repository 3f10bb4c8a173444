#!/bin/bash
# ncu capture of the wide flash attention kernel, then the full GPU suite, smoke and the default + C3 bench lines
mkdir -p gpurun_out
L=gpurun_out/r2_gpu14.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $L
REPS=1 run 600 ncu --set full --clock-control none --import-source on -k regex:attn_wide -c 1 -o gpurun_out/r02_attn_wide python tools/bench_vae_attn.py 16 512
run 2400 python -m pytest tests/ -q -x -m gpu
run 300 python __graft_entry__.py smoke
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$?" >> $L
python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 bench rc=$?" >> $L
