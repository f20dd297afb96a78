#!/bin/bash
# pass L: dense threshold sharing with counters
mkdir -p gpurun_out
for share in 1 0; do
  (HX_DENSE_SHARE=$share HX_DENSE_DBGPRINT=1 timeout -s KILL 200 python bench.py --workload dense --steps 2 --warmup 1 --no-cpu > gpurun_out/r2l_dense_cnt$share.json 2> gpurun_out/r2l_dense_cnt$share.err)
  grep "dense dbg" gpurun_out/r2l_dense_cnt$share.err | tail -2
  (HX_DENSE_SHARE=$share timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2l_dense_share$share.json 2> gpurun_out/r2l_dense_share$share.err)
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_dense_share$share.json').read().strip().split('\n')[-1])
print('share$share', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d.get('recall_at_10_vs_exact_scan'))
PY
done
