#!/bin/bash
# streaming-kernel timings (both precisions), graph A/B with and without the side stream, ncu --set full of norm_act kernels
mkdir -p gpurun_out
T0=$SECONDS
python tools/bench_streaming.py bf16 > gpurun_out/r02g_streaming_bf16.txt 2>&1; cat gpurun_out/r02g_streaming_bf16.txt
python tools/bench_streaming.py fp32 > gpurun_out/r02g_streaming_fp32.txt 2>&1; cat gpurun_out/r02g_streaming_fp32.txt
echo "[t] streaming $((SECONDS-T0)) s"
for ov in 1 0; do
PCRL_OVERLAP_WGRAD=$ov timeout 300 python bench.py --precision bf16 --no-also --no-cpu-baseline --steps 12 > gpurun_out/r02g_bench_bf16_ov$ov.json 2>/dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r02g_bench_bf16_ov$ov.json'))
print("overlap=$ov", {k:d[k] for k in ("value","ms_per_step","eager_step")}, d["clocks"]["sm_mhz"])
PY
done
echo "[t] bench A/B $((SECONDS-T0)) s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:norm_act -c 14 -o gpurun_out/r02g_prof_normact_bf16 -f python tools/bench_streaming.py bf16 "64x64x32 C=64" > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_ncu.log
ls -la gpurun_out/*.ncu-rep
echo "[t] total $((SECONDS-T0)) s"
