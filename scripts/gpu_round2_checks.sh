#!/bin/bash
# Round-2 GPU checks (run under gpurun on one B200): parity tests of the new paths, sanitizers, sub-benches, ncu captures.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_filtered.py tests/test_gpu_service.py -x -q -m gpu --timeout=200 2>&1 | tail -30) > gpurun_out/r2d_filtered_tests.log 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
(timeout 400 $SAN --tool racecheck --racecheck-report analysis --print-limit 10 python scripts/san_driver.py cta scan service seq 2>&1 | grep -v "^=========     and" | tail -30) > gpurun_out/r2d_sanitizer_racecheck.log 2>&1
(timeout 300 $SAN --tool synccheck --num-cuda-barriers 262144 --print-limit 10 python scripts/san_driver.py ring cta policy scan dense 2>&1 | tail -20) > gpurun_out/r2d_sanitizer_synccheck.log 2>&1
(timeout 300 python bench.py --workload dense --steps 10 --warmup 3 > gpurun_out/r2d_bench_dense.json 2> gpurun_out/r2d_bench_dense.err)
(timeout 300 python bench.py --workload prefilter --steps 10 --warmup 3 > gpurun_out/r2d_bench_prefilter.json 2> gpurun_out/r2d_bench_prefilter.err)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dense_scores -s 2 -c 1 -o gpurun_out/r2d_dense python bench.py --workload dense --steps 2 --warmup 1 --no-cpu > gpurun_out/r2d_ncu_dense.log 2>&1)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -o gpurun_out/r2d_scan_topk python bench.py --workload prefilter --steps 2 --warmup 1 --no-cpu > gpurun_out/r2d_ncu_scan.log 2>&1)
(timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hnsw_search|k_validate|k_pack|k_scan|k_select|k_merge|k_simhash" -c 200 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-subresults --no-default-mode > gpurun_out/r2d_launch_bench.log 2>&1)
tail -5 gpurun_out/r2d_filtered_tests.log; tail -4 gpurun_out/r2d_sanitizer_racecheck.log; tail -3 gpurun_out/r2d_sanitizer_synccheck.log; tail -2 gpurun_out/r2d_bench_dense.err gpurun_out/r2d_bench_prefilter.err
