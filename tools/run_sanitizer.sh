#!/bin/bash
# compute-sanitizer memcheck over the op-level tests of the kernels added / changed in round 2:
# the 2-D path's kernels (csrc/planar.cu + the plain GEMMs they feed), the rolling-window convolution and the
# weight-gradient kernel with its new stage shapes (bench-layer shapes, small batch).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_planar_gpu.py -q -m gpu -x -p no:cacheprovider \
  -k "conv2d_im2col or maxpool or add_relu or bilinear or conv_to_3" > gpurun_out/r02z_sanitizer_planar.log 2>&1
echo "sanitizer (2-D ops) exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r02z_sanitizer_planar.log | head -10
echo "[t] $((SECONDS-T0)) s"
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider \
  -k "conv3d_wgrad or conv3d_fp32 or conv3d_fprop" > gpurun_out/r02z_sanitizer_conv3d.log 2>&1
echo "sanitizer (3-D conv ops) exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r02z_sanitizer_conv3d.log | head -10
echo "[t] total $((SECONDS-T0)) s"
