#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_inflate.py -x -q --timeout=100 2>&1 | tail -8 > gpurun_out/r2_c11_t_inflate.log; tail -5 gpurun_out/r2_c11_t_inflate.log
for t in 1 0; do
BIODB_PILEUP_TILE=$t timeout 300 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-cpu --no-extra 2> gpurun_out/r2_c11_bench_tile$t.err | tail -1 > gpurun_out/r2_c11_bench_tile$t.json
echo "== tile $t"; python tools/show_bench.py gpurun_out/r2_c11_bench_tile$t.json || tail -5 gpurun_out/r2_c11_bench_tile$t.err
python -c "
import json; d=json.load(open('gpurun_out/r2_c11_bench_tile$t.json')); e=d['e2e']; print('e2e', round(e['value']/1e6,1), 'ms', round(e['ms_per_step'],1), 'd2h GB', round(e['d2h_bytes_per_step']/1e9,2), e['stage_ms'])"
done
CMD="python bench.py --reads 4000000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate" -c 60 --csv --log-file gpurun_out/launches_r2_inflate.csv $CMD > gpurun_out/launches_r2_inflate.log 2>&1
grep -c inflate gpurun_out/launches_r2_inflate.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"inflate_decode_kernel|inflate_resolve_kernel" -s 4 -c 2 -o gpurun_out/prof_r2_tok $CMD > gpurun_out/prof_r2_tok.log 2>&1
tail -2 gpurun_out/prof_r2_tok.log
timeout 300 python -m pytest tests/test_pileup_chunks.py -x -q --timeout=150 2>&1 | tail -4
