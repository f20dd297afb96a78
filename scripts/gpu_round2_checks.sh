mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu --timeout=300 2>&1 | tail -15) > gpurun_out/r2c_full_gpu_suite.log 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
(timeout 600 $SAN --tool memcheck --print-limit 20 python scripts/san_driver.py 2>&1 | tail -40) > gpurun_out/r2c_sanitizer_memcheck.log 2>&1
(timeout 600 $SAN --tool racecheck --racecheck-report analysis --print-limit 20 python scripts/san_driver.py ring cta scan service seq 2>&1 | tail -60) > gpurun_out/r2c_sanitizer_racecheck.log 2>&1
(timeout 300 $SAN --tool synccheck --print-limit 20 python scripts/san_driver.py ring cta policy scan dense 2>&1 | tail -30) > gpurun_out/r2c_sanitizer_synccheck.log 2>&1
# launch list of one short default bench step (kernel shares), then full captures of the three dominant kernels
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-subresults --no-default-mode > gpurun_out/r2c_launch_bench.log 2>&1)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -o gpurun_out/r2c_scan_topk python bench.py --workload prefilter --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c_ncu_scan.log 2>&1)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_dense_scores -s 2 -c 1 -o gpurun_out/r2c_dense python bench.py --workload dense --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c_ncu_dense.log 2>&1)
tail -4 gpurun_out/r2c_full_gpu_suite.log; tail -5 gpurun_out/r2c_sanitizer_memcheck.log; tail -5 gpurun_out/r2c_sanitizer_racecheck.log; tail -3 gpurun_out/r2c_sanitizer_synccheck.log; ls -la gpurun_out | grep r2c
