#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_mgpu4.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nproc >> $L
run 300 python -m pytest tests/test_pipeline_gpu.py tests/test_b200_worker_gpu.py -q -x -k "bucket or batch or two_real or png or latent"
B200_PNG=gpu run 600 python bench.py --pool-workers 4 --steps 10
B200_PNG=gpu run 600 python bench.py --pool-workers 1 --steps 10
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --config c5 --gpus 4 --steps 3 --warmup 3
run 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/run_sdxl_pp.py --check --peer --iters 2
