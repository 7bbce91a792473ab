#!/bin/bash
# last GPU call of round 2: the core parity tests and the smoke entry on the committed library
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -k "records_match_oracle or pileup_columns_match_oracle or corrupted or synthetic_mixed or stream_stops or truncated" tests/test_gpu_configs.py::test_config0_make_pileup_example -q -m gpu --timeout=200 -p no:cacheprovider > gpurun_out/ch_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/ch_tests.log | cut -c1-400 | tail -8
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
