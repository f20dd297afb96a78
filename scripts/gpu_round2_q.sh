#!/bin/bash
# C5-shaped dense shard (d = 1536, 4096 queries) on one GPU with fewer rows than the 12.5 M of a real C5 shard (time per row is what matters)
mkdir -p gpurun_out
(timeout -s KILL 400 python bench.py --workload dense --dim 1536 --dense-batch 4096 --dense-rows-per-gpu 2000000 --steps 6 --warmup 3 --no-cpu > gpurun_out/r2q_dense_c5shape.json 2> gpurun_out/r2q_dense_c5shape.err)
tail -2 gpurun_out/r2q_dense_c5shape.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_dense_c5shape.json').read().strip().split('\n')[-1])
print('c5 shape', d['value'], d['ms_per_step'], d['roofline'], d.get('recall_at_10_vs_exact_scan'))
PY
