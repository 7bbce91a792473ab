for L in libbiod_b200.so libbiod_b200_v2.so libbiod_b200_v3.so; do echo == $L; BIODB_LIB=$PWD/biod_b200/$L timeout 600 python bench.py --reads 8000000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | tail -1 > /tmp/o.json; python -c "
import json; d=json.load(open('/tmp/o.json')); print(d['value'], d['ms_per_step'], d['inflate_out_gbs'], d['roofline']['stage_ms'])" || tail -c 600 /tmp/o.json; done
