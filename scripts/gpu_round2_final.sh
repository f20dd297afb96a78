#!/bin/bash
# Round-2 final pass on one B200: the whole GPU suite, smoke(), the default bench line, the reference arm, ncu capture of the
# dense kernel and the launch list of the bench command.
mkdir -p gpurun_out
(timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -12) > gpurun_out/r2z_gpu_suite.log 2>&1
(timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2z_smoke.log 2>&1
(timeout -s KILL 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err)
(timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err)
(timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_dense_scores -s 2 -c 1 -o gpurun_out/r2z_dense python bench.py --workload dense --steps 2 --warmup 1 --no-cpu > gpurun_out/r2z_ncu_dense.log 2>&1)
(timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hnsw_search|k_validate|k_pack|k_scan|k_select|k_merge|k_simhash|k_dense|k_filtered" -c 300 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-d1536 > gpurun_out/r2z_launch_bench.log 2>&1)
tail -3 gpurun_out/r2z_gpu_suite.log; cat gpurun_out/r2z_smoke.log; tail -2 gpurun_out/r2z_bench_n1.err; wc -c gpurun_out/r2z_bench_n1.json gpurun_out/r2z_bench_reference.json
