#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/micro/umma_rate > gpurun_out/r02_umma_rate.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_umma_rate.txt
