#!/bin/bash
# pass Z (last GPU seconds): top-T capacity of the dense epilogue, HXD_T = 8 (default) vs 12 vs 16, all with the interleaved
# placement: dense bench (kernel time), dense tests, hazard fixtures (consecutive ids and the adversarial stride-4 layout)
mkdir -p gpurun_out
V=$PWD/helix-db_b200/_variants
for t in 12 16; do
  (HELIX_B200_LIB=$V/libhelix_b200_il1_t$t.so timeout -s KILL 60 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2z_bench_dense_t$t.json 2> gpurun_out/r2z_bench_dense_t$t.err)
  (HELIX_B200_LIB=$V/libhelix_b200_il1_t$t.so timeout -s KILL 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_reference_contracts.py -q -m gpu -k "dense" --timeout=50 2>&1 | tail -3) > gpurun_out/r2z_dense_tests_t$t.log 2>&1
  (HELIX_B200_LIB=$V/libhelix_b200_il1_t$t.so timeout -s KILL 60 python scripts/dense_hazard_check.py 2>&1 | tail -1) > gpurun_out/r2z_hazard_t$t.json 2>&1
  tail -1 gpurun_out/r2z_dense_tests_t$t.log; wc -c gpurun_out/r2z_bench_dense_t$t.json; cut -c1-400 gpurun_out/r2z_hazard_t$t.json
done
(timeout -s KILL 60 python scripts/dense_hazard_check.py 2>&1 | tail -1) > gpurun_out/r2z_hazard_t8.json 2>&1
(timeout -s KILL 60 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2z_bench_dense_t8.json 2> gpurun_out/r2z_bench_dense_t8.err)
cut -c1-400 gpurun_out/r2z_hazard_t8.json; wc -c gpurun_out/r2z_bench_dense_t8.json
