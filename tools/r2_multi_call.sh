#!/bin/bash
# multi-GPU call of round 2: the node's host<->device ceiling with all ranks copying at once, then the bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" > gpurun_out/lscpu_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$PCIE" != "0" ]; then
  timeout 300 $TR --master-port 29511 tools/pcie_bw.py > gpurun_out/pcie_bw_n$N.json 2> gpurun_out/pcie_bw_n$N.err
  tail -c 1500 gpurun_out/pcie_bw_n$N.json
fi
if [ "$2" != "pcie-only" ]; then
  timeout 1500 $TR --master-port 29512 bench.py --gpus $N $3 > gpurun_out/bench_r2_n$N$TAG.json 2> gpurun_out/bench_r2_n$N$TAG.err
  tail -c 400 gpurun_out/bench_r2_n$N$TAG.err
  python - $N$TAG <<'PY'
import json,sys
for l in open('gpurun_out/bench_r2_n%s.json' % sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','inflate_out_gbs')})
        print('e2e',d.get('e2e'))
        print('halo',d['config'].get('halo'), d.get('stitch'))
        print('config4',d.get('config4'))
        print('maq',d.get('maq_e2e'))
PY
fi
