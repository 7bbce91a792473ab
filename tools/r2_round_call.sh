#!/bin/bash
# ncu capture of the MD / MAQ kernels (rows N1, N3)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:"maq_kernel|md_len_kernel|md_replay_kernel" -s 3 -c 3 -o gpurun_out/prof_r2_maq python tools/maq_profile.py 6000000 > gpurun_out/prof_r2_maq.log 2>&1
tail -3 gpurun_out/prof_r2_maq.log
ls -la gpurun_out/prof_r2_maq.ncu-rep
