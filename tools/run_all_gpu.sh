#!/bin/bash
# Full GPU validation: every op-level test group, the whole-model tests, smoke, and the default bench.
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh gemms conv3d_fprop conv3d_dgrad conv3d_wgrad stem convT unshuffled norm_act heads chan1 sgd fp32 > /dev/null
grep -E "^(=====|exit|FAILED|ERROR|E  )|passed|failed" gpurun_out/gpu_tests.log | head -60
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -15
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json | cut -c1-400
