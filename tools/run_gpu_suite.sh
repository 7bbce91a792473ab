#!/bin/bash
# The whole GPU suite with per-test timeouts and the slowest tests listed.  $1 = overall limit in seconds, $2 = extra
# pytest flags (the driver runs it with -x; without it one call shows every failure).
mkdir -p gpurun_out
timeout ${1:-1500} python -m pytest tests -q -m gpu --timeout=400 --durations=15 -p no:cacheprovider $2 > gpurun_out/pytest_gpu_r2_full.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_r2_full.log | tail -40
tail -60 gpurun_out/pytest_gpu_r2_full.log > gpurun_out/pytest_gpu_r2.log
