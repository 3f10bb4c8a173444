#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_mgpu8c.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --config c5 --c5-group 2 --gpus 8 --steps 3 --warmup 3
