#!/bin/bash
# Two-GPU checks: the NCCL path of hx_shard_group (2-rank parity test), then the N=2 bench line (replicas + sharded + dense_c4)
# and the reference arm launched the way the driver launches it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
(timeout -s KILL 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout=400 2>&1 | tail -30) > gpurun_out/r2k_sharded_tests.log 2>&1
(timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err)
tail -5 gpurun_out/r2k_sharded_tests.log; tail -5 gpurun_out/r2k_bench_n2.err; wc -c gpurun_out/r2k_bench_n2.json
