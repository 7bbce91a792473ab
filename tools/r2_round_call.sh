#!/bin/bash
# one GPU call of round 2: the chunk tests (bins.bam now in 85 chunks) + the level-1 leg of the inflate sweep
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pileup_chunks.py -q -m gpu --timeout=200 -p no:cacheprovider > gpurun_out/cf_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/cf_tests.log | cut -c1-400 | tail -8
timeout 200 python tools/inflate_sweep.py --max-gib 8 --levels 1 > gpurun_out/inflate_sweep_r2c.jsonl 2> gpurun_out/inflate_sweep_r2c.err
python - <<'PY'
import json
for l in open('gpurun_out/inflate_sweep_r2c.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('sweep', d['level'], d['gib'], round(d['out_gbs'],1), round(d['algorithmic_gbs'],1), round(d['frac_of_hbm_peak'],4))
PY
