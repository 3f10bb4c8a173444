#!/bin/bash
# wide flash attention for the VAE mid block: kernel tests, timings, pipeline parity, A/B bench, then the default bench
mkdir -p gpurun_out
L=gpurun_out/r2_gpu12.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $L
run 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_wide"
run 300 python tools/bench_vae_attn.py 16 512
run 300 python tools/bench_vae_attn.py 8 768
run 1200 python -m pytest tests/test_pipeline_gpu.py tests/test_sdxl_gpu.py tests/test_patch_parallel_gpu.py -q -x
for f in 0 1 0 1; do
  DL_VAE_FLASH=$f python bench.py --no-pool-e2e --no-cpu-baseline > gpurun_out/r02_ab_vae_flash$f.json 2>> $L; echo "flash=$f rc=$?" >> $L
  python -c "import json;d=json.load(open('gpurun_out/r02_ab_vae_flash$f.json'));print('flash=$f',d['value'],d['ms_per_step'],d['clocks'])" >> $L 2>&1
done
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default bench rc=$?" >> $L
python bench.py --config c3 --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 bench rc=$?" >> $L
