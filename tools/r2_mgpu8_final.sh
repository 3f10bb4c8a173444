#!/bin/bash
# ONE process, WorkerPool with 8 B200Worker threads, requests in -> PNG out (device-side PNG writer)
mkdir -p gpurun_out
timeout 600 python bench.py --pool-workers 8 --steps 40 > gpurun_out/r02_pool_8workers_pipelined.json 2> gpurun_out/r02_pool_8workers_pipelined.err; echo "rc=$?" >> gpurun_out/r02_pool_8workers_pipelined.err
