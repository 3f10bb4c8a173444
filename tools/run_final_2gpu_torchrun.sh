#!/bin/bash
# the driver's scaling launcher at N = 2 on the final code: b200 arm and reference arm
mkdir -p gpurun_out
L=gpurun_out/run_final_2gpu.log
: > $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_torchrun.json 2>> $L; echo "b200 rc=$?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_2gpu_reference.json 2>> $L; echo "ref rc=$?" >> $L
