#!/bin/bash
# inflate parameter variants (rounds per super-chunk, sub-sequence length, match list) against the committed build
mkdir -p gpurun_out
for L in libbiod_b200.so libbiod_b200_r4.so libbiod_b200_r6.so libbiod_b200_s192.so libbiod_b200_m96.so; do
  BIODB_LIB=$PWD/biod_b200/$L timeout 120 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-e2e --no-cpu --no-extra 2>/dev/null | tail -1 > gpurun_out/var_$L.json
  python - $L <<'PY'
import json,sys
d=json.loads(open('gpurun_out/var_%s.json'%sys.argv[1]).read())
print(sys.argv[1], round(d['value']/1e6,1), round(d['ms_per_step'],2), d['roofline']['stage_ms'], d['inflate_counters']['blocks_given_up'], d['checks'])
PY
done
