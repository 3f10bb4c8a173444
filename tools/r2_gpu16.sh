#!/bin/bash
# igemm two-CTA clusters with multicast weight tiles: parity (kernel tests), micro A/B, pipeline parity, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2_gpu16.log
: > $L
run() { echo "### $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "rc=$?" >> $L; }
run 900 python -m pytest tests/test_kernels_gpu.py -q -x
for c in 0 1; do
  echo "## DL_IGEMM_CLUSTER=$c" >> $L
  DL_IGEMM_CLUSTER=$c timeout 600 python tools/bench_igemm.py >> $L 2>&1; echo "rc=$?" >> $L
done
run 1200 python -m pytest tests/test_pipeline_gpu.py tests/test_sdxl_gpu.py -q -x
for c in 0 1 0 1; do
  DL_IGEMM_CLUSTER=$c python bench.py --no-pool-e2e --no-cpu-baseline > gpurun_out/r02_ab_cluster$c.json 2>> $L; echo "cluster=$c rc=$?" >> $L
  python -c "import json;d=json.load(open('gpurun_out/r02_ab_cluster$c.json'));print('cluster=$c',d['value'],d['ms_per_step'],d['clocks'],d['roofline'].get('frac'))" >> $L 2>&1
done
