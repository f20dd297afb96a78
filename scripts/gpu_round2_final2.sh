#!/bin/bash
# Round-2 closing pass: whole GPU suite + smoke + the default bench line with the final kernels.
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -6) > gpurun_out/r2y_gpu_suite.log 2>&1
(timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2) > gpurun_out/r2y_smoke.log 2>&1
(timeout -s KILL 900 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err)
tail -3 gpurun_out/r2y_gpu_suite.log; cat gpurun_out/r2y_smoke.log; tail -2 gpurun_out/r2y_bench_n1.err; wc -c gpurun_out/r2y_bench_n1.json
