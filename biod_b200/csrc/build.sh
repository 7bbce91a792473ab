#!/bin/bash
# Builds biod_b200/libbiod_b200.so (sm_100a only).  No zlib, no CPU fallback in the product.
set -e
cd "$(dirname "$0")"
OUT=../libbiod_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function"
mkdir -p ../_build
for f in inflate records pileup runtime pileup_api; do
  if [ ! -f ../_build/$f.o ] || [ $f.cu -nt ../_build/$f.o ] || [ kernels.h -nt ../_build/$f.o ] || [ pileup.h -nt ../_build/$f.o ] || [ runtime.h -nt ../_build/$f.o ] || [ scan.cuh -nt ../_build/$f.o ] || [ ../../include/biod_b200.h -nt ../_build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o ../_build/$f.o &
  fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT ../_build/inflate.o ../_build/records.o ../_build/pileup.o ../_build/runtime.o ../_build/pileup_api.o -cudart static
echo built $OUT
