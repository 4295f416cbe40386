#!/bin/bash
# 2-D path: op tests, model forward, full step, trainer entry point
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/planar_parity.txt
timeout 900 python -m pytest tests/test_planar_gpu.py -m gpu -q -x --timeout 600 ${PYTEST_ARGS} > gpurun_out/r02r_pytest_planar.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/r02r_pytest_planar.log
