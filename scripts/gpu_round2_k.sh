#!/bin/bash
# Round-2 GPU checks, pass K: golden fixture on the device, dense best-case threshold experiment, build quality (batched vs sequential).
mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests/test_golden.py -q -m gpu --timeout=200 2>&1 | tail -15) > gpurun_out/r2k_golden_tests.log 2>&1
for dbg in 0 8; do
  (HX_DENSE_DEBUG=$dbg timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2k_dense_dbg$dbg.json 2> gpurun_out/r2k_dense_dbg$dbg.err)
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_dense_dbg$dbg.json').read().strip().split('\n')[-1])
print('dbg$dbg', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])
PY
done
(timeout -s KILL 900 python scripts/build_quality.py 200000 > gpurun_out/r2k_build_quality.log 2>&1)
tail -4 gpurun_out/r2k_golden_tests.log; tail -5 gpurun_out/r2k_build_quality.log
