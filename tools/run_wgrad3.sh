#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider -k "conv3d or fp32 or gemm or convT or unshuffled or heads or stem" 2>&1 | tail -8
echo "[t] op tests $((SECONDS-T0)) s"
PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_fp32.txt 2>&1; tail -14 gpurun_out/layers_fp32.txt | cut -c1-45,117-
echo "--- NONPAIR"
PCRL_WGRAD_NONPAIR=1 PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 2>&1 | tail -14 | cut -c1-45,117-
timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_bf16.txt 2>&1; tail -1 gpurun_out/layers_bf16.txt
echo "[t] layers $((SECONDS-T0)) s"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -3
echo "[t] total $((SECONDS-T0)) s"
