#!/bin/bash
# round 2, call F (8 GPUs): DDP equivalence test (2 ranks) + N=8 bench (both precisions, graph step)
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l; nproc
timeout 300 python -m pytest tests/test_ddp_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/r02f_pytest_ddp2.log 2>&1
grep -n "^E \|rel-L2" gpurun_out/r02f_pytest_ddp2.log | head -10; tail -3 gpurun_out/r02f_pytest_ddp2.log
echo "[t] tests $((SECONDS-T0)) s"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err; echo "bench exit $?"; tail -3 gpurun_out/r02f_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_n8.json'))
print({k:d[k] for k in ("value","ms_per_step","host_ms_per_step","replicas_in_sync","eager_step")})
print(d["e2e"]); print({k:d["also"][k] for k in ("value","ms_per_step","host_ms_per_step","replicas_in_sync","eager_step","e2e")})
PY
echo "[t] total $((SECONDS-T0)) s"
