#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:norm_act_bwd -c 8 -o gpurun_out/r02j_prof_normactbwd_bf16 -f python tools/bench_streaming.py bf16 "64x64x32 C=64" > gpurun_out/r02j_ncu.log 2>&1
tail -2 gpurun_out/r02j_ncu.log
