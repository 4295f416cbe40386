"""The 'real competitor' of SURVEY 8(d): the reference's arithmetic (oracle restatement = the same
torch functional ops the reference modules call) run by PyTorch/cuDNN eager on the SAME B200, at the
bench workload (b=32, 64x64x32 + 6x16^3).  Not a bench arm of the driver contract and not product
code: a measurement aid, like the CPU baseline.  Prints one JSON line per mode:
  fp32 (cudnn.allow_tf32 = True, torch's default), fp32 with TF32 off, autocast(bf16).
"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcrlv2_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
batch = orc.synthetic_batch(B, seed=42)
x1, x2, gt = (t.to(dev) for t in batch[:3])
lv = [v.to(dev) for v in batch[3]]


def run(tag, tf32, autocast, channels_last=False, steps=5, warmup=2):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    sd = {k: v.to(dev) for k, v in orc.init_state(0).items()}
    bufs = {}
    rng = random.Random(42)
    times = []
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            orc.train_step(sd, bufs, x1, x2, gt, lv, 0, 1e-3, rng)
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = sorted(times)[len(times) // 2] * 1e3
    print(json.dumps({"impl": "torch-cudnn-eager", "mode": tag, "batch": B, "ms_per_step": ms,
                      "volumes_per_s": B / (ms / 1e3), "torch": torch.__version__,
                      "cudnn": torch.backends.cudnn.version()}), flush=True)


run("fp32 (TF32 convs, torch default)", True, False)
run("autocast bf16", True, True)
run("fp32 (TF32 off)", False, False, steps=3, warmup=1)
