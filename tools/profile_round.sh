#!/bin/bash
# Captures the ncu evidence of one round on the GPU box (run under gpurun):  launch list of the bench command, and
# one `--set full` capture of the hot kernels.  Outputs land in gpurun_out/; tools/summarize_profiles.py turns
# them into the CSV/JSON summaries committed under profiles/.
set -x
R=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --reads 8000000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/launches_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"inflate_decode_kernel|inflate_resolve_kernel|entries_kernel|scan_extract_kernel" -s 8 -c 8 -o gpurun_out/prof_$R $CMD > gpurun_out/prof_$R.log 2>&1
# the column-stationary entries kernel in the plain form (BIODB_PILEUP_TILE=2), for comparison with entries_kernel
BIODB_PILEUP_TILE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"entries_tile_kernel" -s 2 -c 2 -o gpurun_out/prof_${R}_tile $CMD > gpurun_out/prof_${R}_tile.log 2>&1
# the encoder kernel of the write path
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"deflate_warp_kernel" -s 1 -c 1 -o gpurun_out/prof_${R}_deflate python tools/deflate_bench.py --reads 200000 > gpurun_out/prof_${R}_deflate.log 2>&1
ls -la gpurun_out
