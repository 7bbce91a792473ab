#!/bin/bash
# last GPU call of round 2: the final library — smoke entry and the MAQ tests (maq.cu was rebuilt last)
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_maq.py -q -m gpu --timeout=150 -p no:cacheprovider 2>&1 | tail -1
