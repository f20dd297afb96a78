#!/bin/bash
# Eight-GPU check: the N=8 bench line the driver's scaling run will ask for (replicas + id-range shards + C4 dense on 10M rows).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
(timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err)
tail -4 gpurun_out/r2p_bench_n8.err; wc -c gpurun_out/r2p_bench_n8.json
(timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2p_bench_n8_reference.json 2> gpurun_out/r2p_bench_n8_reference.err)
wc -c gpurun_out/r2p_bench_n8_reference.json
