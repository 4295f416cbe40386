#!/bin/bash
# 2 GPUs: DDP equivalence test (with hang diagnostics) + graph tests + N=2 bench (bf16 headline, no also)
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest tests/test_ddp_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/r02h_pytest_ddp2.log 2>&1
grep -n "^E \|rel-L2\|File \"/root/repo" gpurun_out/r02h_pytest_ddp2.log | head -40; tail -3 gpurun_out/r02h_pytest_ddp2.log
echo "[t] ddp test $((SECONDS-T0)) s"
timeout 300 python -m pytest tests/test_graph_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "[t] tests $((SECONDS-T0)) s"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --precision bf16 --no-also > gpurun_out/r02h_bench_n2_bf16.json 2> gpurun_out/r02h_bench_n2.err; echo "bench exit $?"; tail -3 gpurun_out/r02h_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench_n2_bf16.json'))
print({k:d[k] for k in ("value","ms_per_step","host_ms_per_step","replicas_in_sync","eager_step")}); print(d["e2e"])
PY
echo "[t] total $((SECONDS-T0)) s"
