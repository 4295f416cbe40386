#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for ov in 1 0; do
PCRL_OVERLAP_WGRAD=$ov timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ov$ov.json'))
print('overlap=$ov', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'also', d['also']['value'], d['also']['ms_per_step'])
PY
done
