#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_inflate.py -x -q --timeout=100 2>&1 | tail -12 > gpurun_out/r2_c10_t_inflate.log; tail -8 gpurun_out/r2_c10_t_inflate.log
for t in 1 0; do
BIODB_PILEUP_TILE=$t timeout 300 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-cpu --no-extra 2> gpurun_out/r2_c10_bench_tile$t.err | tail -1 > gpurun_out/r2_c10_bench_tile$t.json
echo "== tile $t"; python tools/show_bench.py gpurun_out/r2_c10_bench_tile$t.json || tail -5 gpurun_out/r2_c10_bench_tile$t.err
python -c "
import json; d=json.load(open('gpurun_out/r2_c10_bench_tile$t.json')); e=d['e2e']; print('e2e', round(e['value']/1e6,1), 'ms', round(e['ms_per_step'],1), 'd2h GB', round(e['d2h_bytes_per_step']/1e9,2))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_maq.py tests/test_pileup_chunks.py tests/test_gpu_md.py -x -q --timeout=150 -k "not bins" 2>&1 | tail -12 > gpurun_out/r2_c10_tests.log
tail -12 gpurun_out/r2_c10_tests.log
