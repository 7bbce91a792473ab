#!/bin/bash
# final GPU call of the round: whole suite the way the driver runs it (-x), sanitizer on the new kernels, the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout=400 --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu_r2_final.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2_final.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -p no:cacheprovider --timeout=500 \
  "tests/test_gpu_x_deflate.py::test_device_bytes_equal_the_host_statement" "tests/test_gpu_maq.py::test_hand_computed_columns_on_gpu" \
  "tests/test_gpu_inflate.py::test_valid_streams_of_every_shape_match_zlib" > gpurun_out/sanitizer_r2.txt 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_r2.txt | head -8
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
tail -c 300 gpurun_out/bench_r2_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err
tail -c 600 gpurun_out/bench_r2_reference_arm.json
timeout 300 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu --no-extra --e2e-input file > gpurun_out/bench_r2_file_input.json 2> gpurun_out/bench_r2_file_input.err
python - <<'PY'
import json
for f in ('bench_r2_final','bench_r2_file_input'):
    for l in open('gpurun_out/%s.json'%f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l)
            print(f, {k:d.get(k) for k in ('value','ms_per_step','checks','clocks')})
            e=d.get('e2e') or {}
            print(' e2e', e.get('value'), e.get('ms_per_step'), e.get('input'), e.get('pcie_d2h_gbs'))
            print(' maq', (d.get('maq_e2e') or {}).get('value'), (d.get('maq_e2e') or {}).get('ms_per_step'))
PY
