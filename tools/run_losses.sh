#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider -k "bn1d or linear or cosine or mse or upsample" 2>&1 | tail -15
echo "[t] op tests $((SECONDS-T0)) s"
timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -8
echo "[t] model tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_losses.json 2> gpurun_out/bench_losses.err; echo "bench exit $?"; tail -2 gpurun_out/bench_losses.err
cut -c1-330 gpurun_out/bench_losses.json
echo "[t] total $((SECONDS-T0)) s"
