#!/bin/bash
# ncu capture of the default-mode (policy) kernel and the strict ring kernel with the closing code
mkdir -p gpurun_out
(timeout -s KILL 700 ncu --set full --clock-control none --import-source on -k regex:"k_hnsw_search_policy" -c 10 -o gpurun_out/r2s_policy python bench.py --steps 2 --warmup 1 --no-cpu --no-subresults --no-sharded --no-d1536 > gpurun_out/r2s_ncu_policy.log 2>&1)
tail -3 gpurun_out/r2s_ncu_policy.log | cut -c1-300
ls -la gpurun_out/r2s_policy.ncu-rep
