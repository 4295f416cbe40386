#!/bin/bash
# One GPU call that validates HEAD and collects the judged evidence of round 2 (files gpurun_out/r02z_*):
#   all GPU tests, smoke, default bench (fp32/TF32 headline + bf16 'also', graph step), the CPU reference arm,
#   configs[3] bench line, per-layer tensor-core timings, streaming-kernel timings, the 2-D bench (configs[4] shard)
#   with the torch/cuDNN competitor, the ncu launch list of the bench command and ncu --set full captures of the
#   dominant kernels.  COMPETITOR=1 also re-runs the unmodified reference through PyTorch/cuDNN (3-D).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
T0=$SECONDS
timeout 1800 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02z_pytest_gpu.log 2>&1
grep -n "^E " gpurun_out/r02z_pytest_gpu.log | head -20; tail -4 gpurun_out/r02z_pytest_gpu.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r02z_smoke.log
timeout 900 python bench.py > gpurun_out/r02z_bench_n1_b32.json 2> gpurun_out/r02z_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02z_bench.err
cut -c1-600 gpurun_out/r02z_bench_n1_b32.json
echo "[t] bench $((SECONDS-T0)) s"
timeout 300 python bench.py --impl reference --steps 2 > gpurun_out/r02z_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r02z_bench_reference.json
timeout 600 python bench.py --workload large --no-cpu-baseline --no-also > gpurun_out/r02z_bench_large_config3.json 2> gpurun_out/r02z_bench_large.err; echo "large exit $?"; tail -2 gpurun_out/r02z_bench_large.err; cut -c1-500 gpurun_out/r02z_bench_large_config3.json
echo "[t] large $((SECONDS-T0)) s"
if [ -n "$COMPETITOR" ]; then
timeout 600 python tools/torch_gpu_baseline.py 32 > gpurun_out/r02z_torch_cudnn_reference_same_b200.jsonl 2> gpurun_out/r02z_torch_cudnn.err; cat gpurun_out/r02z_torch_cudnn_reference_same_b200.jsonl
fi
timeout 300 python tools/bench_layers.py 32 > gpurun_out/r02z_layers_bf16.txt 2>&1; tail -1 gpurun_out/r02z_layers_bf16.txt
PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 > gpurun_out/r02z_layers_fp32.txt 2>&1; tail -1 gpurun_out/r02z_layers_fp32.txt
python tools/bench_streaming.py bf16 > gpurun_out/r02z_streaming_bf16.txt 2>&1
python tools/bench_streaming.py fp32 > gpurun_out/r02z_streaming_fp32.txt 2>&1
echo "[t] layers $((SECONDS-T0)) s"
timeout 600 python tools/bench_2d.py > gpurun_out/r02z_bench_2d.jsonl 2> gpurun_out/r02z_bench_2d.err; cut -c1-400 gpurun_out/r02z_bench_2d.jsonl
echo "[t] 2-D $((SECONDS-T0)) s"
if [ -z "$SKIP_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 2800 --csv \
  --log-file gpurun_out/r02z_launches_bench_fp32.csv python bench.py --eager --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r02z_launches_bench_fp32.csv
for prec in fp32 bf16; do
  PCRL_PREC=$prec timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_roll_kernel -s 1 -c 1 \
    -o gpurun_out/r02z_prof_roll_$prec -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_roll_$prec.log 2>&1
  PCRL_PREC=$prec timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor_kernel -s 1 -c 1 \
    -o gpurun_out/r02z_prof_wgrad_$prec -f python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_wgrad_$prec.log 2>&1
done
ls -la gpurun_out/r02z_prof_*.ncu-rep
fi
echo "[t] total $((SECONDS-T0)) s"
