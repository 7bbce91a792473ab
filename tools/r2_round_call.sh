#!/bin/bash
# one GPU call of round 2: the tests touched since the last full run + the streamed-file e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_region.py "tests/test_gpu_parity.py::test_streamed_file_equals_in_memory" -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/cc_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|AssertionError" gpurun_out/cc_tests.log | cut -c1-600 | tail -20
timeout 300 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu --no-extra --e2e-input file > gpurun_out/bench_r2_file_input.json 2> gpurun_out/bench_r2_file_input.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r2_file_input.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); e=d.get('e2e') or {}
        print('file input', d['value'], 'e2e', e.get('value'), e.get('ms_per_step'), e.get('input'), e.get('pcie_d2h_gbs'), e.get('stage_ms'))
PY
