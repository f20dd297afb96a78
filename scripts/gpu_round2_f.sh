#!/bin/bash
# Round-2 GPU checks, pass F: full GPU suite after the first-generation kernels were retired, synccheck with the mbarrier
# rendezvous of the policy CTA build, prefilter bench with the ACORN walk on a built graph.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -40) > gpurun_out/r2f_gpu_suite.log 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
(timeout 500 $SAN --tool synccheck --num-cuda-barriers 262144 --print-limit 6 python scripts/san_driver.py ring cta policy scan dense service 2>&1 | grep -v "Host Frame" | head -120) > gpurun_out/r2f_sanitizer_synccheck.log 2>&1
(timeout 500 $SAN --tool memcheck --print-limit 6 python scripts/san_driver.py ring cta policy service 2>&1 | grep -v "Host Frame" | head -80) > gpurun_out/r2f_sanitizer_memcheck.log 2>&1
(timeout 500 python bench.py --workload prefilter --steps 10 --warmup 3 > gpurun_out/r2f_bench_prefilter.json 2> gpurun_out/r2f_bench_prefilter.err)
tail -12 gpurun_out/r2f_gpu_suite.log; tail -6 gpurun_out/r2f_sanitizer_synccheck.log; tail -4 gpurun_out/r2f_sanitizer_memcheck.log; tail -2 gpurun_out/r2f_bench_prefilter.err
