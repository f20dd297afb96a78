#!/bin/bash
# pass V (closing verification, 8 GPU-minutes left): whole GPU suite incl. the new delete / mode-contract tests, smoke, short C2 line
mkdir -p gpurun_out
(timeout -s KILL 420 python -m pytest tests -q -m gpu --timeout=300 2>&1 | tail -40) > gpurun_out/r2v_gpu_suite.log 2>&1
tail -8 gpurun_out/r2v_gpu_suite.log
(timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2v_smoke.log 2>&1
cat gpurun_out/r2v_smoke.log
(timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-subresults --no-sharded --cpu-seconds 4 > gpurun_out/r2v_bench_c2.json 2> gpurun_out/r2v_bench_c2.err)
tail -2 gpurun_out/r2v_bench_c2.err; wc -c gpurun_out/r2v_bench_c2.json
