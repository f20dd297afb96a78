#!/bin/bash
# pass X: the library with the interleaved dense row placement as its default — whole GPU suite (incl. the id-cluster test)
mkdir -p gpurun_out
(timeout -s KILL 160 python -m pytest tests -q -m gpu --timeout=120 2>&1 | tail -30) > gpurun_out/r2x_gpu_suite.log 2>&1
tail -5 gpurun_out/r2x_gpu_suite.log
