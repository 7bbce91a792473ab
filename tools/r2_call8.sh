#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_inflate.py -x -q 2>&1 | tail -12 > gpurun_out/r2_c8_t_inflate.log; tail -12 gpurun_out/r2_c8_t_inflate.log
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q --durations=12 2>&1 | tail -30 > gpurun_out/r2_c8_parity.log; tail -30 gpurun_out/r2_c8_parity.log
