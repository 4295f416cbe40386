#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_bench_shapes_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x -k "conv3" > gpurun_out/r02m_pytest.log 2>&1; grep -n "^E " gpurun_out/r02m_pytest.log | head; tail -3 gpurun_out/r02m_pytest.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 200 python tools/bench_layers.py 32 2>&1 | tail -15
PCRL_PREC=fp32 timeout 200 python tools/bench_layers.py 32 2>&1 | tail -15
echo "---- NOROLL"
PCRL_IGEMM_NOROLL=1 timeout 200 python tools/bench_layers.py 32 2>&1 | tail -2
PCRL_IGEMM_NOROLL=1 PCRL_PREC=fp32 timeout 200 python tools/bench_layers.py 32 2>&1 | tail -2
echo "[t] total $((SECONDS-T0)) s"
