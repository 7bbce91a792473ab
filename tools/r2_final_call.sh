#!/bin/bash
# final GPU call of the round: whole suite the way the driver runs it (-x), the smoke entry, the inflate sweep, the bench
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -x -q -m gpu --timeout=400 --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu_r2_final.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2_final.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python tools/inflate_sweep.py --max-gib 8 > gpurun_out/inflate_sweep_r2.jsonl 2> gpurun_out/inflate_sweep_r2.err
python - <<'PY'
import json
for l in open('gpurun_out/inflate_sweep_r2.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('sweep', d['level'], d['gib'], round(d['out_gbs'],1), round(d['algorithmic_gbs'],1), round(d['frac_of_hbm_peak'],4))
PY
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
tail -c 300 gpurun_out/bench_r2_final.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r2_final.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','checks')}, d['clocks'].get('samples'))
        e=d.get('e2e') or {}
        print(' e2e', e.get('value'), e.get('ms_per_step'), e.get('pcie_d2h_gbs'))
        print(' maq', (d.get('maq_e2e') or {}).get('value'), 'config3', (d.get('config3') or {}).get('value'))
PY
