#!/bin/bash
# one GPU call of round 2: the tests touched since the last full run, then the ncu evidence (tools/profile_round.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_md.py tests/test_gpu_maq.py "tests/test_gpu_inflate.py::test_fast_path_takes_all_fixture_blocks" tests/test_gpu_x_deflate.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/cb_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|AssertionError" gpurun_out/cb_tests.log | cut -c1-600 | tail -20
timeout 600 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra --md > gpurun_out/bench_r2_md.json 2> gpurun_out/bench_r2_md.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r2_md.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step')}, 'md', d.get('md_reference_bases'))
PY
bash tools/profile_round.sh r2 > gpurun_out/profile_round.log 2>&1
tail -3 gpurun_out/profile_round.log
