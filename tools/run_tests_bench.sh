#!/bin/bash
# Full GPU test-suite + smoke + default bench (short form of run_round.sh, no ncu).
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.err
cut -c1-300 gpurun_out/bench_default.json
echo "[t] total $((SECONDS-T0)) s"
