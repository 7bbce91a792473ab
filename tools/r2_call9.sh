#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_inflate.py -x -q --timeout=100 2>&1 | tail -12 > gpurun_out/r2_c9_t_inflate.log; tail -6 gpurun_out/r2_c9_t_inflate.log
BIODB_INFLATE=tok timeout 300 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-e2e --no-cpu --no-extra 2> gpurun_out/r2_c9_bench_tok.err | tail -1 > gpurun_out/r2_c9_bench_tok.json
python tools/show_bench.py gpurun_out/r2_c9_bench_tok.json || tail -5 gpurun_out/r2_c9_bench_tok.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_maq.py tests/test_pileup_chunks.py -x -q --timeout=150 --durations=8 2>&1 | tail -30 > gpurun_out/r2_c9_tests.log
tail -30 gpurun_out/r2_c9_tests.log
