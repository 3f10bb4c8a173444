#!/bin/bash
# final single-GPU call: ncu launch list (time + DRAM + L2 bytes) of one eager pass -> profiles/r02_ncu_dram_per_kernel.*,
# full GPU suite, smoke, the default bench line (+ reference arm), C3, and the pool with two workers on the one GPU
mkdir -p gpurun_out
L=gpurun_out/r2_final2.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $L; nproc >> $L
run 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/one_pass.py
python tools/ncu_dram_summary.py gpurun_out/r02_launches.csv profiles/r02_ncu_dram_per_kernel > /dev/null 2>> $L; echo "summary rc=$?" >> $L
cp profiles/r02_ncu_dram_per_kernel.json profiles/r02_ncu_dram_per_kernel.txt gpurun_out/
run 2400 python -m pytest tests/ -q -x -m gpu
run 300 python __graft_entry__.py smoke
run 900 python bench.py --impl reference --steps 3 --warmup 1
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$?" >> $L
python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 bench rc=$?" >> $L
python bench.py --pool-workers 1 --steps 10 > gpurun_out/r02_pool_1w_1gpu.json 2>> $L; echo "pool1 rc=$?" >> $L
python bench.py --pool-workers 2 --steps 10 > gpurun_out/r02_pool_2w_1gpu.json 2>> $L; echo "pool2 rc=$?" >> $L
rm -f gpurun_out/r02_launches.csv.gz; gzip -9 -k gpurun_out/r02_launches.csv 2>/dev/null; rm -f gpurun_out/r02_launches.csv
