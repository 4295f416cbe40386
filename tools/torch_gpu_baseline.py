"""The 'real competitor' of SURVEY 8(d) / BASELINE.md section 4 item 4: the UNMODIFIED reference
(`baseline/_ref/train_3d.py:train_pcrlv2_inner` driving `baseline/_ref/models/pcrlv2_model_3d.py:PCRLv23d`)
run by PyTorch/cuDNN eager on the SAME B200, at the bench workload (b=32, 64x64x32 + 6x16^3 views),
with the reference's own `cudnn.benchmark = True` (train_3d.py:58).  Not an arm of the driver contract
and not product code: a measurement aid, like the CPU baseline.

`baseline/_ref/` is a verbatim, git-ignored copy of four reference files made by
`__graft_entry__.build()` in the build container (it travels to the GPU box with gpurun).  The 2-D model
that `models/__init__.py` also imports needs segmentation_models_pytorch (absent): stubbed, never called.
Modes (one JSON line each):
  fp32 (cudnn.allow_tf32 = True: torch's default, what the reference runs), the same with
  channels_last_3d weights/inputs, autocast(bf16) [+ channels_last_3d], fp32 with TF32 off.
Only what the reference itself could be configured to do: no kernels of this repository are involved.
"""
import json
import os
import random
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.exists(os.path.join(REF, "train_3d.py")):
    print(json.dumps({"impl": "torch-cudnn-eager", "unavailable": "baseline/_ref/ is missing: run __graft_entry__.build() "
                      "in the build container (it copies the reference files there)"}))
    sys.exit(0)
import torch

for name in ["segmentation_models_pytorch", "segmentation_models_pytorch.base", "segmentation_models_pytorch.base.modules",
             "segmentation_models_pytorch.base.initialization"]:
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
sys.modules["segmentation_models_pytorch.base"].modules = sys.modules["segmentation_models_pytorch.base.modules"]
sys.modules["segmentation_models_pytorch"].base = sys.modules["segmentation_models_pytorch.base"]
init = sys.modules["segmentation_models_pytorch.base.initialization"]
init.initialize_decoder = init.initialize_head = lambda *a, **k: None
sys.path.insert(0, REF)
import train_3d as ref_train          # noqa: E402  (the reference, unmodified)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
STEPS, WARM = 6, 3
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
VOL, LOC = (64, 64, 32), (16, 16, 16)


def batch(channels_last):
    def t(shape, uniform=False):
        x = torch.rand(shape, generator=g) if uniform else torch.randn(shape, generator=g)
        x = x.to(dev)
        return x.contiguous(memory_format=torch.channels_last_3d) if channels_last else x
    return (t((B, 1) + VOL), t((B, 1) + VOL), t((B, 1) + VOL, True), t((B, 1) + VOL, True),
            [t((B, 1) + LOC) for _ in range(6)])


def run(tag, tf32, autocast, channels_last=False, steps=STEPS, warmup=WARM):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True          # reference train_3d.py:58
    torch.manual_seed(42)
    random.seed(42)
    model = ref_train.PCRLv23d().cuda()
    if channels_last:
        model = model.to(memory_format=torch.channels_last_3d)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    crit, cos = torch.nn.MSELoss().cuda(), torch.nn.CosineSimilarity().cuda()
    args = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    b = batch(channels_last)
    times = []
    devnull = open(os.devnull, "w")
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, sys.stdout = sys.stdout, devnull
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                ref_train.train_pcrlv2_inner(args, 0, [b], model, opt, crit, cos)     # one iteration
        finally:
            sys.stdout = out
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = sorted(times)[len(times) // 2] * 1e3
    print(json.dumps({"impl": "torch-cudnn-eager (unmodified reference trainer + model)", "mode": tag, "batch": B,
                      "ms_per_step": ms, "volumes_per_s": B / (ms / 1e3), "cudnn_benchmark": True,
                      "channels_last_3d": channels_last, "torch": torch.__version__,
                      "cudnn": torch.backends.cudnn.version()}), flush=True)
    del model, opt


run("fp32 (TF32 convs, torch default)", True, False)
run("fp32 (TF32 convs) + channels_last_3d", True, False, channels_last=True)
run("autocast bf16", True, True)
run("autocast bf16 + channels_last_3d", True, True, channels_last=True)
run("fp32 (TF32 off)", False, False, steps=2, warmup=1)
