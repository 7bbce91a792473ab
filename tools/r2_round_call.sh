#!/bin/bash
# the MAQ kernel with its register sorting network: parity tests, then one ncu launch time
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_maq.py -q -m gpu --timeout=150 -p no:cacheprovider > gpurun_out/ck_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/ck_tests.log | cut -c1-400 | tail -6
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size --clock-control none -k regex:"maq_kernel" -s 1 -c 1 --csv --log-file gpurun_out/maq_kernel_time.csv python tools/maq_profile.py 6000000 > gpurun_out/maq_time.log 2>&1
grep -E "maq_kernel" gpurun_out/maq_kernel_time.csv | cut -d, -f5,12-20 | cut -c1-300
