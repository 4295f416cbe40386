#!/bin/bash
# wgrad stage-size experiment (K padding vs pipeline depth) + default bench with clocks/power of the eager/graph A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
{
for nr in "" 3 4; do
  echo "== fp32 PCRL_WGRAD_NROWS=$nr"
  PCRL_PREC=fp32 PCRL_WGRAD_NROWS=$nr timeout 300 python tools/bench_layers.py 32 tr64
done
for nr in "" 6 7 8; do
  echo "== bf16 PCRL_WGRAD_NROWS=$nr"
  PCRL_PREC=bf16 PCRL_WGRAD_NROWS=$nr timeout 300 python tools/bench_layers.py 32 tr64
done
for nr in "" 6 ; do
  echo "== fp32 tr128 PCRL_WGRAD_NROWS=$nr"
  PCRL_PREC=fp32 PCRL_WGRAD_NROWS=$nr timeout 300 python tools/bench_layers.py 32 tr128
done
for nr in "" 12 15; do
  echo "== bf16 tr128 PCRL_WGRAD_NROWS=$nr"
  PCRL_PREC=bf16 PCRL_WGRAD_NROWS=$nr timeout 300 python tools/bench_layers.py 32 tr128
done
} > gpurun_out/r02o_wgrad_nrows.txt 2>&1
tail -70 gpurun_out/r02o_wgrad_nrows.txt
timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
echo "bench exit $?"; tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['eager_step'])
a=d['also']; print(a['value'], a['ms_per_step'], a['clocks'], a.get('eager'))
PY
