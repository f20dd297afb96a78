#!/bin/bash
mkdir -p gpurun_out
(timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu --callers-seconds 0.3 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err)
tail -3 gpurun_out/r2o_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_bench_n1.json').read().strip().split('\n')[-1])
e=d['euclid_d1536']; print('d1536', e['value'], e['e2e'])
print('c2', d['value'], d['e2e']['value'])
for sh in d['prefilter']['reference_shapes']: print(sh['shape'], sh.get('acorn_walk',{}).get('qps'), sh.get('acorn_walk',{}).get('exact_scan_qps_same_queries'))
PY
