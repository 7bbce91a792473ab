#!/bin/bash
# round-2 call 1: correctness of the two-warp inflate kernel + A/B against the one-warp kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_c1_gpu.txt; nproc >> gpurun_out/r2_c1_gpu.txt; free -g >> gpurun_out/r2_c1_gpu.txt; df -h /dev/shm >> gpurun_out/r2_c1_gpu.txt; lscpu | head -20 >> gpurun_out/r2_c1_gpu.txt; numactl -H >> gpurun_out/r2_c1_gpu.txt 2>&1; nvidia-smi topo -m >> gpurun_out/r2_c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_inflate.py -x -q 2>&1 | tail -25 > gpurun_out/r2_c1_t_inflate.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -25 > gpurun_out/r2_c1_t_parity.log
for v in "par libbiod_b200.so" "duo libbiod_b200.so" "duo libbiod_b200_v18.so"; do set -- $v
  BIODB_INFLATE=$1 BIODB_LIB=$PWD/biod_b200/$2 timeout 600 python bench.py --reads 20000000 --steps 3 --warmup 1 --no-e2e --no-cpu 2> gpurun_out/r2_c1_bench_$1_$2.err | tail -1 > gpurun_out/r2_c1_bench_$1_$2.json
  echo "== $1 $2"; python tools/show_bench.py gpurun_out/r2_c1_bench_$1_$2.json || tail -5 gpurun_out/r2_c1_bench_$1_$2.err
done
tail -5 gpurun_out/r2_c1_t_inflate.log; tail -5 gpurun_out/r2_c1_t_parity.log
