#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/model_parity.txt
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/model_tests.log
echo "exit: $?" >> gpurun_out/model_tests.log
timeout 600 python bench.py --steps 4 --warmup 2 --batch ${BENCH_B:-8} --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench exit $?" >> gpurun_out/model_tests.log
tail -5 gpurun_out/bench_small.err >> gpurun_out/model_tests.log
cat gpurun_out/model_tests.log
