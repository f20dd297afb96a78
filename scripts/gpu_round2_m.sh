#!/bin/bash
# pass M: dense epilogue on 16 warps
mkdir -p gpurun_out
(timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "dense" --timeout=200 2>&1 | tail -8) > gpurun_out/r2m_dense_tests.log 2>&1
tail -3 gpurun_out/r2m_dense_tests.log
(timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu > gpurun_out/r2m_dense.json 2> gpurun_out/r2m_dense.err)
python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_dense.json').read().strip().split('\n')[-1])
print('dense', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac'], d.get('recall_at_10_vs_exact_scan'))
PY
(timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu --dense-batch 4096 > gpurun_out/r2m_dense_b4096.json 2> gpurun_out/r2m_dense_b4096.err)
python - <<PY
import json
d=json.loads(open('gpurun_out/r2m_dense_b4096.json').read().strip().split('\n')[-1])
print('dense b4096', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac'], d.get('recall_at_10_vs_exact_scan'))
PY
