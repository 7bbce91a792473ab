#!/bin/bash
# last GPU call of round 2: the freshly rebuilt library — smoke entry, inflate / write-path / region tests
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_x_writer.py tests/test_gpu_x_deflate.py tests/test_gpu_region.py -q -m gpu --timeout=200 -p no:cacheprovider > gpurun_out/cj_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/cj_tests.log | cut -c1-400 | tail -8
