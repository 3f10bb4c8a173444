#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gpu10.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sdxl_gpu.py tests/test_b200_worker_gpu.py tests/test_fp32_mode_gpu.py -q -x -s -k "not 768"
DL_UNET_LN_FOLD=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
DL_ATTN_X=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
DL_UNET_LN_FOLD=0 DL_UNET_GN_FUSE=0 DL_ATTN_X=0 DL_ATTN_PP=0 run 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-pool-e2e
