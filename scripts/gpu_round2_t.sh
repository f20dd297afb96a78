#!/bin/bash
# pass T: relaxed launch bounds (CTA policy build, QCH = 48 ring build): whole suite + the C2 line with the default-mode object
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -6) > gpurun_out/r2t_gpu_suite.log 2>&1
tail -3 gpurun_out/r2t_gpu_suite.log
(timeout -s KILL 500 python bench.py --steps 5 --warmup 3 --no-subresults --no-sharded --cpu-seconds 4 > gpurun_out/r2t_bench_c2.json 2> gpurun_out/r2t_bench_c2.err)
tail -2 gpurun_out/r2t_bench_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench_c2.json').read().strip().split('\n')[-1])
dm=d['default_mode']
print('c2', d['value'], d['roofline']['frac'], 'policy', dm['kernel_qps'], dm['alg_GBps'], 'single', dm['single_query_us'], dm.get('cpu_port_identical_to_device'), d['clocks'], 'b1', d['single_stream_batch1'])
PY
