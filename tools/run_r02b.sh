#!/bin/bash
# round 2, GPU call B: captured-graph step tests, the step tests with the measured bounds, default bench
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_graph_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/r02b_pytest_graph.log
tail -40 gpurun_out/r02b_pytest_graph.log
echo "[t] graph tests $((SECONDS-T0)) s"
timeout 1500 python -m pytest tests/test_step_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02b_pytest_step.log
tail -15 gpurun_out/r02b_pytest_step.log
echo "[t] step tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02b_bench.err
cut -c1-3000 gpurun_out/r02b_bench.json
echo "[t] total $((SECONDS-T0)) s"
