#!/bin/bash
# Round-2 GPU checks, pass J: whole suite after the graph_dirty fix; dense: TMEM loads without the column tests.
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -30) > gpurun_out/r2j_gpu_suite.log 2>&1
for dbg in 0 4; do
  (HX_DENSE_DEBUG=$dbg timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2j_dense_dbg$dbg.json 2> gpurun_out/r2j_dense_dbg$dbg.err)
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_dense_dbg$dbg.json').read().strip().split('\n')[-1])
print('dbg$dbg', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])
PY
done
tail -5 gpurun_out/r2j_gpu_suite.log
