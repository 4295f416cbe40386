#!/bin/bash
# compute-sanitizer memcheck over the op-level tests of the kernels changed this round
mkdir -p gpurun_out
T0=$SECONDS
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider \
  -k "conv3d_wgrad or conv3d_fp32 or norm_act_wide or bn1d or linear or cosine or mse or upsample or convT" \
  > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|Error" gpurun_out/sanitizer.log | head -20
echo "[t] total $((SECONDS-T0)) s"
