#!/bin/bash
# Round-2 GPU checks, pass H: dense limiter experiments (epilogue / TMA switched off), policy kernel after the paired reductions.
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  (HX_DENSE_DEBUG=$dbg timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2h_dense_dbg$dbg.json 2> gpurun_out/r2h_dense_dbg$dbg.err)
done
(timeout -s KILL 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_scale.py -q -m gpu -x --timeout=500 2>&1 | tail -15) > gpurun_out/r2h_policy_tests.log 2>&1
(timeout -s KILL 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-subresults --no-sharded --no-d1536 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err)
for dbg in 0 1 2 3; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_dense_dbg$dbg.json').read().strip().split('\n')[-1])
    print('dbg$dbg', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])
except Exception as e: print('dbg$dbg failed', e)
PY
done
tail -5 gpurun_out/r2h_policy_tests.log; tail -2 gpurun_out/r2h_bench_c2.err
