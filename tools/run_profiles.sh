#!/bin/bash
# ncu evidence for the bench command (profiles/): launch list of ~1.5 steps of the default bench
# and one --set full capture of the dominant kernel at its largest launch shape.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2600 --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_bench.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kmajor -s 1 -c 2 \
  -o gpurun_out/prof_kmajor_top python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_mnmajor -s 1 -c 1 \
  -o gpurun_out/prof_mnmajor_top python tools/bench_layers.py 32 up_tr64.ops.0 > gpurun_out/ncu_top2.log 2>&1
ls -la gpurun_out/*.ncu-rep
