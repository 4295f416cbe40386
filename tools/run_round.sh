#!/bin/bash
# One GPU call that validates HEAD and collects the judged evidence:
#   op-level + whole-model GPU tests, smoke, default bench (fp32/TF32 headline + bf16 'also'),
#   the CPU reference arm, per-layer tensor-core timings, and ncu captures for profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
T0=$SECONDS
timeout 1200 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.err
cut -c1-1500 gpurun_out/bench_default.json
echo "[t] bench $((SECONDS-T0)) s"
timeout 300 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_bf16.txt 2>&1; tail -14 gpurun_out/layers_bf16.txt
PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 > gpurun_out/layers_fp32.txt 2>&1; tail -14 gpurun_out/layers_fp32.txt
echo "[t] layers $((SECONDS-T0)) s"
if [ -z "$SKIP_NCU" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kmajor -s 1 -c 2 \
  -o gpurun_out/prof_kmajor_bf16 -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top.log 2>&1
PCRL_PREC=fp32 timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kmajor -s 1 -c 2 \
  -o gpurun_out/prof_kmajor_fp32 -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top_fp32.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_bf16 -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top2.log 2>&1
PCRL_PREC=fp32 timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_fp32 -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top3.log 2>&1
echo "[t] ncu full $((SECONDS-T0)) s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2600 --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_bench.csv
ls -la gpurun_out/*.ncu-rep
fi
echo "[t] total $((SECONDS-T0)) s"
