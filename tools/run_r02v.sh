#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/planar_parity.txt
timeout 900 python -m pytest tests/test_planar_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/r02v_pytest_planar.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r02v_pytest_planar.log
timeout 800 python tools/bench_2d.py --skip_torch > gpurun_out/r02v_bench_2d.jsonl 2> gpurun_out/r02v_bench_2d.err; tail -3 gpurun_out/r02v_bench_2d.err; cut -c1-700 gpurun_out/r02v_bench_2d.jsonl
