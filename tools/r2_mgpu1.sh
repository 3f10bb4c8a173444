#!/bin/bash
# 2-GPU call: multi-worker pool test, pool bench (one process, 2 workers), torchrun c2 + c5, C5 teacher-forced check
mkdir -p gpurun_out
L=gpurun_out/r2_mgpu1.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi -L >> $L
run 300 python -m pytest tests/test_b200_worker_gpu.py -q -x -k "two_real_workers or gpu_png"
B200_PNG=gpu run 400 python bench.py --pool-workers 2 --steps 10
B200_PNG=pil run 400 python bench.py --pool-workers 2 --steps 10
B200_PNG=gpu run 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config c5 --gpus 2 --steps 3 --warmup 3
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/run_sdxl_pp.py --check --peer --iters 2
