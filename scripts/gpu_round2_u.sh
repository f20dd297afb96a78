#!/bin/bash
mkdir -p gpurun_out
(timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "dense" --timeout=250 2>&1 | tail -12) > gpurun_out/r2u_dense_tests.log 2>&1
tail -5 gpurun_out/r2u_dense_tests.log
