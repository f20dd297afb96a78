#!/bin/bash
# Round-2 GPU checks, pass E: ACORN classify rewrite, policy named barrier, synccheck, prefilter bench with the walk.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_filtered.py tests/test_gpu_policy.py tests/test_gpu_service.py -x -q -m gpu --timeout=300 2>&1 | tail -15) > gpurun_out/r2e_tests.log 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
(timeout 400 $SAN --tool synccheck --num-cuda-barriers 262144 --print-limit 10 python scripts/san_driver.py ring cta policy scan dense service 2>&1 | tail -20) > gpurun_out/r2e_sanitizer_synccheck.log 2>&1
(timeout 400 python bench.py --workload prefilter --steps 10 --warmup 3 > gpurun_out/r2e_bench_prefilter.json 2> gpurun_out/r2e_bench_prefilter.err)
(timeout 300 python scripts/service_sweep.py > gpurun_out/r2e_service_sweep.log 2>&1)
tail -6 gpurun_out/r2e_tests.log; tail -5 gpurun_out/r2e_sanitizer_synccheck.log; tail -2 gpurun_out/r2e_bench_prefilter.err; tail -12 gpurun_out/r2e_service_sweep.log
