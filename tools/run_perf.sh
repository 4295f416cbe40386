#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_layers.py ${1:-32} > gpurun_out/layers.txt 2>&1
cat gpurun_out/layers.txt
