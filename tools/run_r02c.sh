#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
python tools/diag_overlap.py 2>&1 | tail -40
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "side_stream" > gpurun_out/r02c_side.log 2>&1; tail -5 gpurun_out/r02c_side.log
timeout 900 python -m pytest tests/test_graph_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02c_pytest_graph.log 2>&1
tail -40 gpurun_out/r02c_pytest_graph.log
echo "[t] graph tests $((SECONDS-T0)) s"
timeout 1500 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "full_step" > gpurun_out/r02c_pytest_step.log 2>&1
grep -n "^E \|assert\|Error" gpurun_out/r02c_pytest_step.log | head -40
tail -5 gpurun_out/r02c_pytest_step.log
echo "[t] total $((SECONDS-T0)) s"
