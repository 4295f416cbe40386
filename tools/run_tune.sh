#!/bin/bash
# kmajor tile-shape sweep (env overrides) per precision
mkdir -p gpurun_out
for prec in fp32 bf16; do
for cfg in "0 0" "1 2" "2 1" "1 4" "2 4" "1 1"; do
  set -- $cfg
  echo "=== $prec MT=$1 P=$2"
  PCRL_PREC=$prec PCRL_IGEMM_MT=$1 PCRL_IGEMM_P=$2 timeout 120 python tools/bench_layers.py 32 2>&1 | tail -14 | cut -c1-88
done
done
