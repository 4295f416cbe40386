#!/bin/bash
# Builds libpcrl_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=${PCRL_OUT:-../libpcrl_b200.so}
BUILD=${PCRL_BUILD_DIR:-build}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $PCRL_EXTRA_FLAGS"
# --use_fast_math only where no transcendental / division decides a result: the tensor-core kernels
# (their arithmetic is the MMA).  The streaming / head / loss kernels use IEEE division, sqrt and expf.
FAST="--use_fast_math"
mkdir -p $BUILD
pids=()
for f in api igemm_kmajor igemm_mnmajor streaming heads losses augment planar; do
  if [ ! -f $BUILD/$f.o ] || [ $f.cu -nt $BUILD/$f.o ] || [ sm100.cuh -nt $BUILD/$f.o ] || [ common.cuh -nt $BUILD/$f.o ]; then
    EXTRA=""; case $f in igemm_*) EXTRA=$FAST;; esac
    nvcc $FLAGS $EXTRA -c $f.cu -o $BUILD/$f.o > $BUILD/$f.log 2>&1 &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { cat $BUILD/*.log | grep -E "error|Error" ; exit 1; }; done
nvcc -shared -o $OUT $BUILD/api.o $BUILD/igemm_kmajor.o $BUILD/igemm_mnmajor.o $BUILD/streaming.o $BUILD/heads.o $BUILD/losses.o $BUILD/augment.o $BUILD/planar.o -lcudart
echo "built $(realpath $OUT)"
