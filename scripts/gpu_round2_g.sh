#!/bin/bash
# Round-2 GPU checks, pass G: CTA-pair dense kernel (tests, bench, ncu), the concurrent-streams failure under memcheck.
mkdir -p gpurun_out
(timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "dense" --timeout=200 2>&1 | tail -25) > gpurun_out/r2g_dense_tests.log 2>&1
(timeout -s KILL 300 python -m pytest tests/test_gpu_service.py tests/test_gpu_sharded.py -q -m gpu -k "concurrent_streams or group_of_one" --timeout=200 2>&1 | tail -25) > gpurun_out/r2g_streams_alone.log 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
(timeout -s KILL 600 $SAN --tool memcheck --print-limit 4 python -m pytest tests/test_gpu_service.py -q -m gpu -k "concurrent_streams" --timeout=500 2>&1 | grep -v "Host Frame" | head -80) > gpurun_out/r2g_streams_memcheck.log 2>&1
if grep -q "passed" gpurun_out/r2g_dense_tests.log && ! grep -q "failed" gpurun_out/r2g_dense_tests.log; then
  (timeout -s KILL 400 python bench.py --workload dense --steps 10 --warmup 3 > gpurun_out/r2g_bench_dense.json 2> gpurun_out/r2g_bench_dense.err)
  (timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_dense_scores -s 2 -c 1 -o gpurun_out/r2g_dense python bench.py --workload dense --steps 2 --warmup 1 --no-cpu > gpurun_out/r2g_ncu_dense.log 2>&1)
  (timeout -s KILL 300 $SAN --tool memcheck --print-limit 4 python scripts/san_driver.py dense 2>&1 | grep -v "Host Frame" | tail -12) > gpurun_out/r2g_dense_memcheck.log 2>&1
fi
tail -8 gpurun_out/r2g_dense_tests.log; tail -6 gpurun_out/r2g_streams_alone.log; tail -30 gpurun_out/r2g_streams_memcheck.log; tail -3 gpurun_out/r2g_bench_dense.err 2>/dev/null; tail -3 gpurun_out/r2g_dense_memcheck.log 2>/dev/null
