#!/bin/bash
# ncu --set full capture of the inflate kernel selected by BIODB_INFLATE (default duo): 2 launches, source view
mkdir -p gpurun_out
K=${1:-inflate_duo_kernel}; TAG=${2:-duo}
CMD="python bench.py --reads 4000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 2 -o gpurun_out/prof_r2_$TAG $CMD > gpurun_out/prof_r2_$TAG.log 2>&1
tail -3 gpurun_out/prof_r2_$TAG.log
