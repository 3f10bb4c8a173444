#!/bin/bash
# ONE process, WorkerPool with 8 B200Worker threads, requests in -> PNG out (device-side PNG writer), with the
# per-batch timing records (device time vs host enqueue time per worker)
mkdir -p gpurun_out
B200_BATCH_TIMING=1 timeout 600 python bench.py --pool-workers 8 --steps 20 > gpurun_out/r02_pool_8workers_timing.json 2> gpurun_out/r02_pool_8workers_timing.err; echo "rc=$?" >> gpurun_out/r02_pool_8workers_timing.err
