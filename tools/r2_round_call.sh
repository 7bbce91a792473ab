#!/bin/bash
# one GPU call of round 2: the sharded tests (spans of shards are new)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -k "sharded" tests/test_gpu_md.py tests/test_gpu_configs.py::test_config3_prefix_in_8_shards -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/cd_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/cd_tests.log | cut -c1-600 | tail -20
