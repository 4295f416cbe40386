#!/bin/bash
# Builds libpcrl_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libpcrl_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math -Xptxas -v"
mkdir -p build
pids=()
for f in api igemm_kmajor igemm_mnmajor streaming heads; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ sm100.cuh -nt build/$f.o ] || [ common.cuh -nt build/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { cat build/*.log | grep -E "error|Error" ; exit 1; }; done
nvcc -shared -o $OUT build/api.o build/igemm_kmajor.o build/igemm_mnmajor.o build/streaming.o build/heads.o -lcudart
echo "built $(realpath $OUT)"
