#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inflate.py -x -q 2>&1 | tail -8 > gpurun_out/r2_c7_t_inflate.log; tail -3 gpurun_out/r2_c7_t_inflate.log
for v in "tok" "par" "duo"; do
  BIODB_INFLATE=$v timeout 600 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-e2e --no-cpu --no-extra 2> gpurun_out/r2_c7_bench_$v.err | tail -1 > gpurun_out/r2_c7_bench_$v.json
  echo "== $v"; python tools/show_bench.py gpurun_out/r2_c7_bench_$v.json || tail -5 gpurun_out/r2_c7_bench_$v.err
done
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_pileup_chunks.py tests/test_gpu_configs.py -x -q 2>&1 | tail -15 > gpurun_out/r2_c7_tests.log
tail -15 gpurun_out/r2_c7_tests.log
