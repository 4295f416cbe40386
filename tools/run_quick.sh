#!/bin/bash
# quick check: norm/act + model tests, then the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider -k "norm_act" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['value'], d['ms_per_step'], 'also', d['also']['value'], d['also']['ms_per_step'])
for k,v in d['kernel_breakdown_ms'].items():
    if v['ms']>0.9: print(f"{k:32s} {v['ms']:8.3f} ms  n={v['n']:4d}  {v.get('tflops','')}")
PY
