#!/bin/bash
# round 2, call E (2 GPUs): DDP equivalence test, N=2 bench; plus graph tests re-check
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_ddp_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/r02e_pytest_ddp2.log 2>&1
grep -n "^E \|Error\|rel-L2" gpurun_out/r02e_pytest_ddp2.log | head -30; tail -4 gpurun_out/r02e_pytest_ddp2.log
timeout 300 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "leaky" 2>&1 | tail -4
echo "[t] tests $((SECONDS-T0)) s"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench exit $?"; tail -3 gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_n2.json'))
print({k:d[k] for k in ("value","ms_per_step","host_ms_per_step","replicas_in_sync","eager_step")})
print(d["e2e"]); print({k:d["also"][k] for k in ("value","ms_per_step","host_ms_per_step","replicas_in_sync","eager_step")})
PY
echo "[t] total $((SECONDS-T0)) s"
