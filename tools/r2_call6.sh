#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inflate.py -x -q 2>&1 | tail -5 > gpurun_out/r2_c6_t_inflate.log
for v in "duo libbiod_b200_vnb.so" "duo libbiod_b200.so" "duo libbiod_b200_v18.so"; do set -- $v
  BIODB_INFLATE=$1 BIODB_LIB=$PWD/biod_b200/$2 timeout 600 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-e2e --no-cpu --no-extra 2> gpurun_out/r2_c6_bench_$1_$2.err | tail -1 > gpurun_out/r2_c6_bench_$1_$2.json
  echo "== $1 $2"; python tools/show_bench.py gpurun_out/r2_c6_bench_$1_$2.json || tail -5 gpurun_out/r2_c6_bench_$1_$2.err
done
tail -3 gpurun_out/r2_c6_t_inflate.log
