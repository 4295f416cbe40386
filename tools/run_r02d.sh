#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 2400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02d_pytest_gpu.log 2>&1
grep -n "^E \|Error" gpurun_out/r02d_pytest_gpu.log | head -40
tail -12 gpurun_out/r02d_pytest_gpu.log
echo "[t] total $((SECONDS-T0)) s"
