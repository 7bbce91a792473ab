#!/bin/bash
# Builds biod_b200/libbiod_b200.so (sm_100a only).  No zlib, no CPU fallback in the product.
set -e
cd "$(dirname "$0")"
OUT=${BIODB_OUT:-../libbiod_b200.so}
BUILD=${BIODB_BUILD:-../_build}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${BIODB_DEFS} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function"
mkdir -p $BUILD
for f in inflate inflate_tok crc32 records region pileup maq mdtag deflate runtime pileup_api; do
  if [ ! -f $BUILD/$f.o ] || [ $f.cu -nt $BUILD/$f.o ] || [ kernels.h -nt $BUILD/$f.o ] || [ pileup.h -nt $BUILD/$f.o ] || [ runtime.h -nt $BUILD/$f.o ] || [ scan.cuh -nt $BUILD/$f.o ] || [ md_chain.h -nt $BUILD/$f.o ] || [ md_walk.h -nt $BUILD/$f.o ] || [ bai.h -nt $BUILD/$f.o ] || [ bai_build.h -nt $BUILD/$f.o ] || [ deflate_enc.h -nt $BUILD/$f.o ] || [ maq.h -nt $BUILD/$f.o ] || [ inflate_common.cuh -nt $BUILD/$f.o ] || [ ../../include/biod_b200.h -nt $BUILD/$f.o ]; then
    rm -f $BUILD/$f.o                      # a failed compile must not leave a stale object behind for the link
    $NVCC $FLAGS -c $f.cu -o $BUILD/$f.o &
    PIDS="$PIDS $!"
  fi
done
for p in $PIDS; do wait $p || { echo "build.sh: a compile failed" >&2; exit 1; }; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $BUILD/inflate.o $BUILD/inflate_tok.o $BUILD/crc32.o $BUILD/records.o $BUILD/region.o $BUILD/pileup.o $BUILD/maq.o $BUILD/mdtag.o $BUILD/deflate.o $BUILD/runtime.o $BUILD/pileup_api.o -cudart static
echo built $OUT
