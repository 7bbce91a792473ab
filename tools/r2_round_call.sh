#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x_deflate.py tests/test_gpu_x_writer.py tests/test_zz_gpu_create_index.py "tests/test_gpu_inflate.py::test_fast_path_takes_all_fixture_blocks" tests/test_pileup_chunks.py -q -m gpu --timeout=200 -p no:cacheprovider > gpurun_out/ca_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|gave_up|AssertionError" gpurun_out/ca_tests.log | cut -c1-600 | tail -20
timeout 300 python tools/deflate_bench.py > gpurun_out/deflate_bench_r2.json 2> gpurun_out/deflate_bench_r2.err
tail -c 1500 gpurun_out/deflate_bench_r2.json; tail -3 gpurun_out/deflate_bench_r2.err
timeout 1200 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 600 gpurun_out/bench_r2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','records_per_sec','inflate_out_gbs')})
        print('e2e',d.get('e2e'))
        print('stage',d['roofline'].get('stage_ms'))
        print('cpu',d.get('cpu_baseline'))
        print('config3',d.get('config3'))
        print('maq',d.get('maq_e2e'))
        print('md',d.get('md_reference_bases'))
PY
