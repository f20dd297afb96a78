#!/bin/bash
# pass R: policy kernel without its own top-k tracker (the beam's prefix is the tracker): parity + speed
mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests/test_gpu_policy.py tests/test_golden.py tests/test_gpu_scale.py -q -m gpu --timeout=500 2>&1 | tail -6) > gpurun_out/r2r_policy_tests.log 2>&1
tail -3 gpurun_out/r2r_policy_tests.log
(timeout -s KILL 500 python bench.py --steps 5 --warmup 3 --no-subresults --no-sharded --no-d1536 --cpu-seconds 4 > gpurun_out/r2r_bench_c2.json 2> gpurun_out/r2r_bench_c2.err)
tail -2 gpurun_out/r2r_bench_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_c2.json').read().strip().split('\n')[-1])
dm=d['default_mode']
print('c2', d['value'], d['roofline']['frac'], 'policy', dm['kernel_qps'], dm['alg_GBps'], dm['single_query_us'], dm.get('cpu_port_identical_to_device'), d['clocks'])
PY
