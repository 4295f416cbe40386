#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_staging_gpu.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02k_pytest_staging.log 2>&1; grep -n "^E " gpurun_out/r02k_pytest_staging.log | head -20; tail -3 gpurun_out/r02k_pytest_staging.log
echo "[t] staging tests $((SECONDS-T0)) s"
python - <<'PY'
# staging throughput: items/s of the GPU augmentation chain (8 volumes per item) at b=32
import time, torch, sys
sys.path.insert(0, '.')
from pcrlv2_b200 import staging as S
b=32
raw=[(torch.rand(b,1,64,64,32).pin_memory(), torch.rand(b,1,64,64,32).pin_memory(), [torch.rand(b,1,16,16,16).pin_memory() for _ in range(6)]) for _ in range(12)]
aug=S.GpuAugmenter("cuda", seed=0)
for _ in S.PrefetchLoader(raw[:3], aug): pass
torch.cuda.synchronize(); t0=time.perf_counter()
n=0
for batch in S.PrefetchLoader(raw, aug): n+=b
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("staging: %.0f items/s (%.2f ms per batch of %d: H2D 53.5 MB + flip/blur/noise/gamma/swap/znorm of 8 volumes per item)" % (n/dt, dt/len(raw)*1e3, b))
PY
echo "[t] total $((SECONDS-T0)) s"
