#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02n_pytest_gpu.log 2>&1
grep -n "^E " gpurun_out/r02n_pytest_gpu.log | head -20; tail -4 gpurun_out/r02n_pytest_gpu.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02n_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_bench.json'))
print({k:d[k] for k in ("value","ms_per_step","eager_step")}, d["e2e"]["value"], d["roofline"]["frac"])
a=d["also"]; print({k:a[k] for k in ("value","ms_per_step","eager_step")}, a["e2e"]["value"], a["roofline"]["frac"])
for k in ("conv3d_k3_fprop","conv3d_k3_dgrad","conv3d_k3_dgrad_unshuffled","conv3d_k3_wgrad","norm_act_bwd","norm_act_fwd"): print(k, d["kernel_breakdown_ms"][k], a["kernel_breakdown_ms"][k])
PY
echo "[t] total $((SECONDS-T0)) s"
