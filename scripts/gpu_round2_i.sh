#!/bin/bash
# Round-2 GPU checks, pass I: service file alone / whole suite (order-dependent failure), dense epilogue prefilter, policy kernel.
mkdir -p gpurun_out
(timeout -s KILL 400 python -m pytest tests/test_gpu_service.py -q -m gpu --timeout=300 2>&1 | tail -15) > gpurun_out/r2i_service_file.log 2>&1
(timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -60) > gpurun_out/r2i_gpu_suite.log 2>&1
(HX_DENSE_DEBUG=0 timeout -s KILL 200 python bench.py --workload dense --steps 10 --warmup 3 > gpurun_out/r2i_dense.json 2> gpurun_out/r2i_dense.err)
(timeout -s KILL 500 python bench.py --steps 5 --warmup 3 --no-cpu --no-subresults --no-sharded --no-d1536 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err)
tail -4 gpurun_out/r2i_service_file.log; tail -8 gpurun_out/r2i_gpu_suite.log
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_dense.json').read().strip().split('\n')[-1])
print('dense', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac'], d.get('recall_at_10_vs_exact_scan'))
d=json.loads(open('gpurun_out/r2i_bench_c2.json').read().strip().split('\n')[-1])
print('c2', d['value'], d['roofline']['frac'], 'policy', d['default_mode']['kernel_qps'], d['default_mode']['alg_GBps'], d['default_mode']['single_query_us'])
PY
