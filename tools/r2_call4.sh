#!/bin/bash
mkdir -p gpurun_out
BIODB_LIB=$PWD/biod_b200/libbiod_b200_vt.so timeout 600 python tools/duo_timing.py 8000000 > gpurun_out/r2_c4_timing.json 2> gpurun_out/r2_c4_timing.err
cat gpurun_out/r2_c4_timing.json; tail -3 gpurun_out/r2_c4_timing.err
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_md.py tests/test_gpu_region.py -x -q 2>&1 | tail -30 > gpurun_out/r2_c4_tests.log
tail -30 gpurun_out/r2_c4_tests.log
