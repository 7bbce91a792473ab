#!/bin/bash
# one GPU call of round 2: inflate tests + the level-0 leg of the inflate sweep after the word-wide stored copy
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_inflate.py "tests/test_gpu_parity.py::test_synthetic_mixed_cigar" -q -m gpu --timeout=120 -p no:cacheprovider > gpurun_out/cg_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/cg_tests.log | cut -c1-400 | tail -8
timeout 200 python tools/inflate_sweep.py --max-gib 4 --levels 0 > gpurun_out/inflate_sweep_r2d.jsonl 2> gpurun_out/inflate_sweep_r2d.err
python - <<'PY'
import json
for l in open('gpurun_out/inflate_sweep_r2d.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('sweep', d['level'], d['gib'], round(d['out_gbs'],1), round(d['algorithmic_gbs'],1), round(d['frac_of_hbm_peak'],4))
PY
