#!/bin/bash
# wgrad experiments: parity of the weight-gradient paths, then per-layer timings with / without PAIR mode.
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider -k "wgrad or fp32 or gemm or convT or unshuffled or heads" 2>&1 | tail -8
echo "[t] op tests $((SECONDS-T0)) s"
timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_bf16.txt 2>&1; tail -14 gpurun_out/layers_bf16.txt
PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_fp32.txt 2>&1; tail -14 gpurun_out/layers_fp32.txt
echo "--- NOPAIR"
PCRL_WGRAD_NOPAIR=1 timeout 300 python tools/bench_layers.py 32 up_tr64 2>&1 | tail -3
PCRL_WGRAD_NOPAIR=1 PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 up_tr64 2>&1 | tail -3
echo "[t] layers $((SECONDS-T0)) s"
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "[t] model tests $((SECONDS-T0)) s"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_bf16_pair -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_w1.log 2>&1
PCRL_PREC=fp32 timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_fp32_pair -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_w2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_bf16_up128 -f python tools/bench_layers.py 32 up_tr128.ops.0 > gpurun_out/ncu_w3.log 2>&1
echo "[t] ncu $((SECONDS-T0)) s"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_wgrad.json 2> gpurun_out/bench_wgrad.err; echo "bench exit $?"
cut -c1-400 gpurun_out/bench_wgrad.json
echo "[t] total $((SECONDS-T0)) s"
