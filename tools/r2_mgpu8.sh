#!/bin/bash
# 8-GPU call: the product's multi-GPU path (ONE process, WorkerPool with 8 B200Workers, requests in -> PNG out),
# with the device-side PNG writer and with the reference's PIL encoder; then config C5 over 8 ranks.
mkdir -p gpurun_out
L=gpurun_out/r2_mgpu8.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi -L >> $L; nproc >> $L
B200_PNG=gpu run 600 python bench.py --pool-workers 8 --steps 10
B200_PNG=pil B200_PNG_THREADS=64 run 400 python bench.py --pool-workers 8 --steps 6
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --config c5 --gpus 8 --steps 3 --warmup 3
