#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/planar_parity.txt
timeout 900 python -m pytest tests/test_planar_gpu.py -m gpu -q -x --timeout 600 -k "graphed" > gpurun_out/r02t_pytest_planar.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r02t_pytest_planar.log; grep "2d graph" gpurun_out/planar_parity.txt
true
