#!/bin/bash
# round 2, GPU call A: new parity tests (not -x: every failure is wanted), whole suite, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
nproc; free -g | head -2
T0=$SECONDS
timeout 1500 python -m pytest tests/test_step_gpu.py tests/test_bench_shapes_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r02a_pytest_new.log
tail -30 gpurun_out/r02a_pytest_new.log
echo "[t] new tests $((SECONDS-T0)) s"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02a_pytest_old.log
tail -8 gpurun_out/r02a_pytest_old.log
echo "[t] old tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02a_bench.err
cut -c1-2500 gpurun_out/r02a_bench.json
echo "[t] total $((SECONDS-T0)) s"
