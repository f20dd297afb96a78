#!/bin/bash
# pass W (closing verification + one A/B, 8 GPU-minutes left).  Priority order: whole GPU suite, smoke, short C2 line; then the
# dense row-placement A/B (HXD_INTERLEAVE): hazard check on both builds, dense tests and the dense bench on the variant.
mkdir -p gpurun_out
(timeout -s KILL 420 python -m pytest tests -q -m gpu --timeout=300 2>&1 | tail -40) > gpurun_out/r2w_gpu_suite.log 2>&1
tail -6 gpurun_out/r2w_gpu_suite.log
(timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2w_smoke.log 2>&1
cat gpurun_out/r2w_smoke.log
VAR=$PWD/helix-db_b200/_variants/libhelix_b200_il1.so
(timeout -s KILL 120 python scripts/dense_hazard_check.py 2>&1 | tail -2) > gpurun_out/r2w_hazard_default.json 2>&1
(HELIX_B200_LIB=$VAR timeout -s KILL 120 python scripts/dense_hazard_check.py 2>&1 | tail -2) > gpurun_out/r2w_hazard_interleave.json 2>&1
cat gpurun_out/r2w_hazard_default.json gpurun_out/r2w_hazard_interleave.json
(HELIX_B200_LIB=$VAR timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -k "dense" --timeout=180 2>&1 | tail -6) > gpurun_out/r2w_dense_tests_interleave.log 2>&1
tail -3 gpurun_out/r2w_dense_tests_interleave.log
(timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-subresults --no-sharded --cpu-seconds 4 > gpurun_out/r2w_bench_c2.json 2> gpurun_out/r2w_bench_c2.err)
tail -2 gpurun_out/r2w_bench_c2.err; wc -c gpurun_out/r2w_bench_c2.json
(HELIX_B200_LIB=$VAR timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2w_bench_dense_interleave.json 2> gpurun_out/r2w_bench_dense_interleave.err)
tail -2 gpurun_out/r2w_bench_dense_interleave.err; wc -c gpurun_out/r2w_bench_dense_interleave.json
(timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2w_bench_dense_default.json 2> gpurun_out/r2w_bench_dense_default.err)
tail -2 gpurun_out/r2w_bench_dense_default.err; wc -c gpurun_out/r2w_bench_dense_default.json
