#!/bin/bash
# final single-GPU call of round 2: full GPU suite, smoke, the reference arm, the default bench line, C3
# (the ncu launch list profiles/r02_ncu_dram_per_kernel.* comes from the same script's earlier run with
#  `ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
#   --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/one_pass.py` + tools/ncu_dram_summary.py)
mkdir -p gpurun_out
L=gpurun_out/run_final_1gpu.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $L; nproc >> $L
run 2400 python -m pytest tests/ -q -x -m gpu
run 300 python __graft_entry__.py smoke
run 900 python bench.py --impl reference --steps 3 --warmup 1
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$?" >> $L
python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 bench rc=$?" >> $L
