#!/bin/bash
# Two-GPU checks: the NCCL path of hx_shard_group (2-rank parity test), then the N=2 bench line (replicas + sharded + dense_c4).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
(timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout=400 2>&1 | tail -30) > gpurun_out/r2e_sharded_tests.log 2>&1
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err)
SAN=/usr/local/cuda/bin/compute-sanitizer
(CUDA_VISIBLE_DEVICES=0 timeout 300 $SAN --tool synccheck --num-cuda-barriers 262144 --print-limit 6 python scripts/san_driver.py policy 2>&1 | head -60) > gpurun_out/r2e_sanitizer_synccheck_policy.log 2>&1
tail -5 gpurun_out/r2e_sharded_tests.log; tail -5 gpurun_out/r2e_bench_n2.err; wc -c gpurun_out/r2e_bench_n2.json
