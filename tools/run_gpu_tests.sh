#!/bin/bash
# Runs the op-level GPU tests group by group, each in its own process with a timeout, so that a
# sticky CUDA error or a hung kernel in one group cannot hide the results of the others.
mkdir -p gpurun_out
L=gpurun_out/gpu_tests.log
: > $L
for k in "$@"; do
  echo "=================== $k" >> $L
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$k" --tb=short -p no:cacheprovider 2>&1 | tail -60 >> $L
  echo "exit: ${PIPESTATUS[0]}" >> $L
done
grep -E "^(=====|exit|FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed" $L | head -150
