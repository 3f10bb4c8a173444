#!/bin/bash
# igemm residual in the epilogue for deep-K layers: parity, micro A/B, pipeline parity, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2_gpu17.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 900 python -m pytest tests/test_kernels_gpu.py -q -x -k igemm
for c in 0 16; do
  echo "## DL_IGEMM_RES_EPI_MIN_KB=$c" >> $L
  DL_IGEMM_RES_EPI_MIN_KB=$c timeout 600 python tools/bench_igemm.py res >> $L 2>&1
  DL_IGEMM_RES_EPI_MIN_KB=$c timeout 600 python tools/bench_igemm.py vae >> $L 2>&1
  DL_IGEMM_RES_EPI_MIN_KB=$c timeout 600 python tools/bench_igemm.py conv >> $L 2>&1; echo "rc=$?" >> $L
done
run 1200 python -m pytest tests/test_pipeline_gpu.py tests/test_sdxl_gpu.py -q -x
for c in 0 16 0 16; do
  DL_IGEMM_RES_EPI_MIN_KB=$c python bench.py --no-pool-e2e --no-cpu-baseline > gpurun_out/r02_ab_resepi$c.json 2>> $L; echo "resepi=$c rc=$?" >> $L
  python -c "import json;d=json.load(open('gpurun_out/r02_ab_resepi$c.json'));print('res_epi_min_kb=$c',d['value'],d['ms_per_step'],d['clocks'],d['roofline'].get('frac'),d['roofline']['by_kernel_ms']['igemm'])" >> $L 2>&1
done
