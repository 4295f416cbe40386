#!/bin/bash
# streaming-kernel timings (both precisions) + ncu --set full of norm_act_bwd / fwd at the full-res shape
mkdir -p gpurun_out
python tools/bench_streaming.py bf16 > gpurun_out/r02g_streaming_bf16.txt 2>&1; cat gpurun_out/r02g_streaming_bf16.txt
python tools/bench_streaming.py fp32 > gpurun_out/r02g_streaming_fp32.txt 2>&1; cat gpurun_out/r02g_streaming_fp32.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:norm_act -c 14 -o gpurun_out/r02g_prof_normact_bf16 -f python tools/bench_streaming.py bf16 "64x64x32 C=64" > gpurun_out/r02g_ncu.log 2>&1
tail -3 gpurun_out/r02g_ncu.log
ls -la gpurun_out/*.ncu-rep
