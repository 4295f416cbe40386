#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r02i_pytest.log 2>&1; grep -n "^E " gpurun_out/r02i_pytest.log | head; tail -3 gpurun_out/r02i_pytest.log
echo "[t] tests $((SECONDS-T0)) s"
python tools/bench_streaming.py bf16 bwd > gpurun_out/r02i_streaming_bf16.txt 2>&1; cat gpurun_out/r02i_streaming_bf16.txt
python tools/bench_streaming.py bf16 fwd | grep fwd
python tools/bench_streaming.py fp32 bwd > gpurun_out/r02i_streaming_fp32.txt 2>&1; cat gpurun_out/r02i_streaming_fp32.txt
python tools/bench_streaming.py fp32 fwd | grep fwd
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02i_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench.json'))
print({k:d[k] for k in ("value","ms_per_step","eager_step")}, d["e2e"]["value"])
print({k:d["also"][k] for k in ("value","ms_per_step","eager_step")}, d["also"]["e2e"]["value"])
for k in ("norm_act_bwd","norm_act_fwd"): print(k, d["kernel_breakdown_ms"][k], d["also"]["kernel_breakdown_ms"][k])
PY
echo "[t] total $((SECONDS-T0)) s"
