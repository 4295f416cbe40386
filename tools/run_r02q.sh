#!/bin/bash
# new wgrad stage policy: full GPU suite, per-layer table, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02q_pytest_gpu.log
PCRL_PREC=fp32 timeout 300 python tools/bench_layers.py 32 > gpurun_out/r02q_layers_fp32.txt 2>&1; tail -1 gpurun_out/r02q_layers_fp32.txt
PCRL_PREC=bf16 timeout 300 python tools/bench_layers.py 32 > gpurun_out/r02q_layers_bf16.txt 2>&1; tail -1 gpurun_out/r02q_layers_bf16.txt
timeout 900 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['eager_step'])
print({k:(v['ms'],v.get('frac')) for k,v in d['kernel_breakdown_ms'].items() if v['ms']>1})
a=d['also']; print(a['value'], a['ms_per_step'], a['e2e']['value'], a['clocks'], a.get('eager'))
print({k:(v['ms'],v.get('frac')) for k,v in a['kernel_breakdown_ms'].items() if v['ms']>1})
PY
