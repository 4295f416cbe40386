"""configs[4] (2-D path) on one B200: per-GPU shard b=8 of the b=64 / 8-GPU configuration, two global
3x224x224 crops + six local 3x96x96 crops per item, synthetic data, random-init weights.

Prints one JSON line per precision: images/s of the full pre-training step (three forwards, loss, backward, SGD)
of this library (captured CUDA-graph step by default, --eager for Python launches), the host cost of a step, the
end-to-end number through train_2d.train_pcrlv2_inner with pinned host batches, and -- the competitor on the same
GPU -- the SAME network written with torch's own ops (the oracle's functional restatement of the reference model
run on CUDA tensors = PyTorch/cuDNN eager with cudnn.benchmark, TF32 allowed as by default; bf16 via autocast).
The reference's own 2-D module cannot be constructed here (segmentation_models_pytorch is absent)."""
import argparse
import json
import os
import random
import statistics
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcrlv2_b200 import _lib
from pcrlv2_b200 import train_2d as T2
from pcrlv2_b200.models import PCRLv2
from pcrlv2_b200.train_3d import FlatSGD


import torch.distributed as dist

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def barrier():
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps):
    """ms per step: CUDA events on this rank's stream between two barriers, max over the ranks."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if WORLD > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def batch(bsz, seed, size, local, dev=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s, rand=False: (torch.rand(s, generator=g) if rand else torch.randn(s, generator=g))
    x1, x2, gt = mk(bsz, 3, *size), mk(bsz, 3, *size), mk(bsz, 3, *size, rand=True)
    lv = [mk(bsz, 3, *local) for _ in range(6)]
    if dev is not None:
        return x1.to(dev), x2.to(dev), gt.to(dev), [v.to(dev) for v in lv]
    if pin:
        return x1.pin_memory(), x2.pin_memory(), gt.pin_memory(), [v.pin_memory() for v in lv]
    return x1, x2, gt, lv


def ours(precision, args):
    dev = torch.device("cuda", LOCAL)
    torch.manual_seed(0)
    m = PCRLv2(precision=precision).to(dev).train()
    if WORLD > 1:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=0)
    opt = FlatSGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    nb = 3
    data = [batch(args.batch, 42 + 100 * RANK + i, args.size, args.local, dev) for i in range(nb)]
    random.seed(42)          # the same draws on every rank

    def eager_step(i):
        x1, x2, gt, lv = data[i % nb]
        loss, *_ = T2.pcrlv2_step_loss(m, x1, x2, gt, lv, 0, crit, cos)
        opt.zero_grad()
        loss.backward()
        opt.step()

    gs = None
    if not args.eager:
        gs = T2.graphed_step_for(m, opt, crit, cos, data[0][0], data[0][3]).capture_all()

    def graph_step(i):
        x1, x2, gt, lv = data[i % nb]
        gs.load(x1, x2, gt, lv)          # device -> device into the graph's static buffers
        gs.run(0, skip_guard=False)

    step = eager_step if gs is None else graph_step
    for i in range(args.warmup):
        step(i)
    _lib.launch_count[0] = 0
    ms = timed(step, args.steps)
    launches = (gs.launches if gs is not None else _lib.launch_count[0] / args.steps)
    host = []
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(i)
        host.append((time.perf_counter() - t0) * 1e3)
    # end to end: pinned host batches -> H2D every step, loss scalars read back (the public trainer call)
    hb = [batch(args.batch, 42 + 100 * RANK + i, args.size, args.local, pin=True) for i in range(nb)]
    k2 = max(3, args.steps // 2)
    loader = [(hb[i % nb][0], hb[i % nb][1], hb[i % nb][2], hb[i % nb][2], hb[i % nb][3]) for i in range(k2)]
    targs = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=precision == "bf16", epochs=240)
    so = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        barrier()
        t0 = time.perf_counter()
        T2.train_pcrlv2_inner(targs, 0, loader, m, opt, crit, cos)
        barrier()
        t1 = time.perf_counter()
    finally:
        sys.stdout = so
    et = torch.tensor([t1 - t0], device=dev)
    sync = None
    if WORLD > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        chk = torch.stack([opt._flat_p.double().sum(), (opt._flat_p.double() ** 2).sum()])
        allc = [torch.zeros_like(chk) for _ in range(WORLD)]
        dist.all_gather(allc, chk)
        sync = bool(all(torch.equal(allc[0], c) for c in allc))
    t0, t1 = 0.0, et.item()
    h2d = sum(t.numel() * 4 for t in hb[0][:3]) + sum(t.numel() * 4 for t in hb[0][3])
    eager = None
    if gs is not None:
        for i in range(2):
            eager_step(i)
        eager = {"ms_per_step": timed(eager_step, max(3, args.steps // 2))}
    if gs is not None:
        opt.__dict__.get("_graphed", {}).clear()       # captured NCCL work must go before the communicator does
    return {"ms_per_step": ms, "value": WORLD * args.batch / ms * 1e3, "host_ms_per_step": statistics.median(host),
            "replicas_in_sync": sync,
            "step": "eager (Python launches)" if gs is None else "CUDA graph replay", "eager_step": eager,
            "gpu_launches_per_step": launches,
            "e2e": {"value": WORLD * args.batch * k2 / (t1 - t0), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 16, "steps": k2},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}


def torch_eager(precision, args):
    """The same network and step with torch's own CUDA ops (cuDNN / cuBLAS / ATen)."""
    from oracle import pcrlv2_oracle_2d as orc
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in orc.init_state(0).items()}
    keys = [k for k in sd if orc.is_param(k)]
    for k in keys:
        sd[k].requires_grad_(True)
    opt = torch.optim.SGD([sd[k] for k in keys], lr=1e-3, momentum=0.9, weight_decay=1e-4)
    nb = 3
    data = [batch(args.batch, 42 + i, args.size, args.local, dev) for i in range(nb)]
    rng = random.Random(42)
    chl = args.channels_last

    def cl(t):
        return t.contiguous(memory_format=torch.channels_last) if chl else t

    def step(i):
        x1, x2, gt, lv = data[i % nb]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16"):
            loss, _, _ = orc.step_loss(sd, cl(x1), cl(x2), gt, [cl(v) for v in lv], 0, rng)
        opt.zero_grad()
        loss.backward()
        opt.step()

    for i in range(max(args.warmup, 4)):
        step(i)
    ms = timed(step, args.steps)
    return {"ms_per_step": ms, "value": args.batch / ms * 1e3}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, nargs=2, default=(224, 224))
    ap.add_argument("--local", type=int, nargs=2, default=(96, 96))
    ap.add_argument("--channels_last", action="store_true")
    ap.add_argument("--skip_torch", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the eager step instead of the captured graph")
    ap.add_argument("--precision", default="both", choices=["both", "fp32", "bf16"])
    args = ap.parse_args(argv)
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
        args.skip_torch = True
    for precision in (("fp32", "bf16") if args.precision == "both" else (args.precision,)):
        r = ours(precision, args)
        line = {"metric": "ChestX-ray 2D pretrain images/sec (configs[4]: b=%d per GPU x %d GPU(s))" % (args.batch, WORLD),
                "value": r["value"], "unit": "images/s", "n_gpus": WORLD, "steps": args.steps, "warmup": args.warmup,
                "replicas_in_sync": r["replicas_in_sync"],
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if precision == "fp32" else "bf16", "data": "synthetic",
                "config": {"workload": "NIH ChestX-ray 2D pretrain ResNet-18 UNet, b=%d/GPU, 2 x 3x%dx%d + 6 x 3x%dx%d, %s"
                           % (args.batch, *args.size, *args.local, precision), "step": r["step"]},
                "eager_step": r["eager_step"],
                "host_ms_per_step": r["host_ms_per_step"], "gpu_launches": r["gpu_launches_per_step"] * args.steps,
                "launches_per_step": r["gpu_launches_per_step"], "e2e": r["e2e"], "peak_mem_gb": r["peak_mem_gb"]}
        if not args.skip_torch:
            try:
                t = torch_eager(precision, args)
                line["torch_cudnn_eager_same_gpu"] = dict(t, channels_last=args.channels_last,
                                                          note="the oracle's functional restatement of the reference model on CUDA tensors")
            except Exception as e:   # noqa: BLE001
                line["torch_cudnn_eager_same_gpu"] = {"error": repr(e)[:200]}
        if RANK == 0:
            print(json.dumps(line), flush=True)
    if WORLD > 1:
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
