#!/bin/bash
# one GPU call of round 2: the write path after the streaming writer (biodb_writer_drain)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_x_writer.py tests/test_gpu_x_deflate.py tests/test_zz_gpu_create_index.py -q -m gpu --timeout=200 -p no:cacheprovider > gpurun_out/ci_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/ci_tests.log | cut -c1-400 | tail -8
