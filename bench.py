#!/usr/bin/env python
"""Benchmark of the PCRLv2 3-D pre-training hot path on B200 (contract: see DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this implementation
  python bench.py --impl reference [...]                          # CPU arm (oracle port of the reference)
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one iteration of train_pcrlv2_inner (reference train_3d.py:109-151) on one synthetic
LUNA-shaped batch: 2 global 64x64x32 forwards + 1 forward over 6 local 16^3 views per sample,
loss, backward, SGD.  metric = samples ("volumes") per second, whole job.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

print_line = print   # main() swaps in a collector so that stdout carries only the JSON line
METRIC = "LUNA 64x64x32 pretrain volumes/sec"
UNIT = "volumes/s"
VOL = (64, 64, 32)
LOCAL = (16, 16, 16)


def flops_per_sample(vol=None, local=None, n_local=6):
    """Algorithmic FLOPs (2*MACs of every conv / convT / linear) of one sample's step
    (SURVEY 8d): forward F(V) scales with the voxel count; backward = 2x forward minus the data
    gradient of the Cin=1 stem."""
    def fwd(v):
        convs = [(1, 32, 1), (32, 64, 1), (64, 64, 8), (64, 128, 8), (128, 128, 64), (128, 256, 64),
                 (256, 256, 512), (256, 512, 512), (512, 256, 64), (256, 256, 64), (256, 128, 8),
                 (128, 128, 8), (128, 64, 1), (64, 64, 1)]
        f = sum(2.0 * (v / s) * 27 * ci * co for ci, co, s in convs)
        f += sum(2.0 * (v / s) * 27 * c for c, s in [(256, 64), (128, 8), (64, 1)])      # ds heads
        f += 2.0 * v * 64                                                                  # 1x1x1
        f += sum(2.0 * (v / s) * ci * co * 8 for ci, co, s in [(512, 512, 512), (256, 256, 64), (128, 128, 8)])
        f += sum(2.0 * (c * 2 * c) * 2 for c in (256, 128, 64))                           # predictor
        return f
    vol, local = vol or VOL, local or LOCAL
    vg = vol[0] * vol[1] * vol[2]
    vl = local[0] * local[1] * local[2]
    f_fwd = 2 * fwd(vg) + n_local * fwd(vl)
    stem_dgrad = (2 * vg + n_local * vl) * 2.0 * 27 * 32
    return f_fwd + 2 * f_fwd - stem_dgrad


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of this rank's GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def measure_tf32_peak(dev, seconds=1.0):
    """cuBLAS TF32 GEMM throughput on this GPU, measured the way MEASURED_PEAKS.json measures the
    bf16 figure (8192^3 matmul, back to back for about `seconds`: a sustained number)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters, total_ms, total_it = 10, 0.0, 0
        while total_ms < seconds * 1e3:
            e0.record()
            for _ in range(iters):
                torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            total_it += iters
        return 2.0 * n ** 3 * total_it / (total_ms / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def set_workload(name):
    """configs[1]/[2] geometry (default) or configs[3]: 128x128x64 crops with 32^3 local views
    (the reference leaves the local size of that config open; 32^3 keeps the 1/32 voxel ratio)."""
    global VOL, LOCAL
    if name == "large":
        VOL, LOCAL = (128, 128, 64), (32, 32, 32)
        for k in list(WORKLOADS):
            WORKLOADS[k] = ("LUNA 3D pretrain 128x128x64 crops + 6x32^3 local views, b={B}/GPU, %s (configs[3]: "
                            "128x128x64 b=8 bf16 on 1xB200, the large-volume regime)" %
                            ("bf16" if k == "bf16" else "fp32 storage / TF32 operands"))


WORKLOADS = {
    "fp32": "LUNA 3D pretrain 64x64x32 + 6x16^3 local views, b={B}/GPU, fp32 storage / TF32 tensor-core "
            "operands (configs[1]: b=32 fp32 on 1xB200; the reference's own fp32 convs run TF32 under "
            "torch's default cudnn.allow_tf32)",
    "bf16": "LUNA 3D pretrain 64x64x32 + 6x16^3 local views, b={B}/GPU, bf16 "
            "(per-GPU shard of configs[2]: b=256 bf16 on 8xB200)",
}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(bsz=2, steps=2, warmup=1):
    """Times the CPU oracle (port of the reference step) on the host cores.  Returns
    (samples/s, cores, description)."""
    from oracle import pcrlv2_oracle as orc
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sd = orc.init_state(0)
    bufs = {}
    rng = random.Random(42)
    batch = orc.synthetic_batch(bsz, seed=42)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.train_step(sd, bufs, batch[0], batch[1], batch[2], batch[3], 0, 1e-3, rng)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    ms = statistics.median(times) * 1e3
    return bsz / (ms / 1e3), cores, ms, f"{warmup} warm-up + {steps} timed CPU steps of batch {bsz} (fp32, torch CPU ops, {cores} threads)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    rate, cores, ms, sample = cpu_oracle_rate(2, steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS["fp32"].format(B=32) + " -- CPU arm: bounded sample of batch-2 steps "
                               "(configs[0]), fp32 torch CPU ops on all host cores"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def make_batch(B, seed, pinned):
    g = torch.Generator().manual_seed(seed)

    def t(shape, uniform=False):
        x = torch.rand(shape, generator=g) if uniform else torch.randn(shape, generator=g)
        return x.pin_memory() if pinned else x
    return (t((B, 1) + VOL), t((B, 1) + VOL), t((B, 1) + VOL, True), t((B, 1) + VOL, True),
            [t((B, 1) + LOCAL) for _ in range(6)])


def measure(args, precision, host, rank, world, dev):
    """One precision mode, every rank: K device-resident steps (the product path = the captured-graph
    step unless --eager), the end-to-end leg through train_pcrlv2_inner, the eager step beside it, and
    one instrumented eager step for the per-kernel breakdown.  Returns a dict (rank 0 prints)."""
    import torch.distributed as dist
    from pcrlv2_b200 import _lib
    from pcrlv2_b200 import train_3d as T
    from pcrlv2_b200.models import PCRLv23d

    torch.manual_seed(42)
    random.seed(42)
    B = args.batch
    nbatches = len(host)
    model = PCRLv23d(precision=precision).to(dev).train()
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
    opt = T.FlatSGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    resident = [(b[0].to(dev), b[1].to(dev), b[2].to(dev), [v.to(dev) for v in b[4]]) for b in host]

    def eager_step(i):
        x1, x2, gt, lv = resident[i % nbatches]
        loss, _, _, _ = T.pcrlv2_step_loss(model, x1, x2, gt, lv, 0, crit, cos)
        opt.zero_grad()
        loss.backward()
        opt.step()

    gs = None
    if not args.eager:
        gs = T.graphed_step_for(model, opt, crit, cos, resident[0][0], resident[0][3])   # captures here
        gs.capture_all()        # the graphs of all three index2 draws, before anything is timed

    def graph_step(i):
        x1, x2, gt, lv = resident[i % nbatches]
        gs.load(x1, x2, gt, lv)          # device -> static input buffers of the graph (53.5 MB D2D)
        gs.run(0)

    device_step = eager_step if gs is None else graph_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def host_cost(step_fn):
        # host time of one step with an EMPTY launch queue (sync before, none inside the bracket):
        # when it approaches ms_per_step the job is host-bound
        ts = []
        for i in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step_fn(i)
            ts.append((time.perf_counter() - t0) * 1e3)
        barrier()
        h = torch.tensor([statistics.median(ts)], device=dev)
        if world > 1:
            dist.all_reduce(h, op=dist.ReduceOp.MAX)
        return h.item()

    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    _lib.launch_count[0] = 0
    ms_total = timed(device_step, args.steps)
    clocks = sampler.stop()
    launches = gs.launches * args.steps if gs is not None else _lib.launch_count[0]
    out = {"value": world * B * args.steps / (ms_total / 1e3), "ms_per_step": ms_total / args.steps,
           "gpu_launches": launches, "launches_per_step": launches / args.steps, "clocks": clocks,
           "precision": precision, "host_ms_per_step": host_cost(device_step),
           "step": "eager (Python launches)" if gs is None else "CUDA graph replay (one cudaGraphLaunch per step)"}

    # ---- end to end through the public trainer call: pinned host batches -> device copies every step,
    # loss scalars read back every step
    import types
    targs = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    k2 = max(3, args.steps // 2)
    loader = [host[i % nbatches] for i in range(k2)]
    barrier()
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        if args.eager:
            os.environ["PCRL_GRAPH"] = "0"
        t0 = time.perf_counter()
        T.train_pcrlv2_inner(targs, 0, loader, model, opt, crit, cos)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    finally:
        sys.stdout = _stdout
    e2e_t = torch.tensor([t1 - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    h2d = sum(t.numel() * 4 for t in (host[0][0], host[0][1], host[0][2])) + sum(v.numel() * 4 for v in host[0][4])
    out["e2e"] = {"value": world * B * k2 / e2e_t.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": 16, "steps": k2}

    # ---- replicas in sync: after all those steps every rank must hold the same parameters (the
    # multi-GPU correctness evidence: same draws, same all-reduced gradients, same update everywhere)
    if world > 1:
        chk = torch.stack([opt._flat_p.double().sum(), (opt._flat_p.double() ** 2).sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        out["replicas_in_sync"] = bool(all(torch.equal(allc[0], c) for c in allc))
        out["param_checksum"] = [float(v) for v in allc[0]]

    # ---- the eager step beside the graph replay (same kernels, launched from Python)
    if gs is not None:
        for i in range(2):
            eager_step(i)
        ke = max(6, args.steps // 2)
        smp = ClockSampler(dev.index)
        smp.start()
        ms_e = timed(eager_step, ke) / ke
        ck_e = smp.stop()
        out["eager"] = {"ms_per_step": ms_e, "host_ms_per_step": host_cost(eager_step), "steps": ke,
                        "sm_mhz": ck_e.get("sm_mhz"), "power_w": ck_e.get("power_w")}
        # the graph step again, right after the eager one (same thermal / power state): what an A/B of
        # the two step kinds should be read from, with the clocks and board power each ran at
        graph_step(0)
        smp = ClockSampler(dev.index)
        smp.start()
        out["eager"]["graph_ms_per_step_measured_right_after"] = timed(graph_step, ke) / ke
        ck_g = smp.stop()
        out["eager"]["graph_sm_mhz"] = ck_g.get("sm_mhz")
        out["eager"]["graph_power_w"] = ck_g.get("power_w")

    # ---- per-kernel roofline: one instrumented EAGER step (events around every entry point).  Every
    # rank runs it (the step contains the gradient all-reduce); only rank 0 reports.
    # The weight-gradient kernels normally run on a side stream, concurrently with main-stream
    # kernels; event-bracketed durations would then include the time a kernel waits for SMs held by
    # the other stream.  The instrumented step therefore runs serialised (PCRL_OVERLAP_WGRAD=0): the
    # per-kernel durations are clean, `value` above was measured with the overlap on.
    torch.cuda.synchronize()
    prev_ov = os.environ.get("PCRL_OVERLAP_WGRAD")
    os.environ["PCRL_OVERLAP_WGRAD"] = "0"
    _lib.profile[0] = []
    try:
        eager_step(0)
        torch.cuda.synchronize()
    finally:
        prof, _lib.profile[0] = _lib.profile[0], None
        if prev_ov is None:
            del os.environ["PCRL_OVERLAP_WGRAD"]
        else:
            os.environ["PCRL_OVERLAP_WGRAD"] = prev_ov
    barrier()
    per = {}
    for name, ints, a, b in prof:
        d = per.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
        d["ms"] += a.elapsed_time(b)
        d["n"] += 1
        if name in ("pcrl_conv3d_k3_fprop", "pcrl_conv3d_k3_dgrad", "pcrl_conv3d_k3_wgrad",
                    "pcrl_conv3d_k3_dgrad_unshuffled"):
            n_, d_, h_, w_, ci, co = ints[-7:-1]       # (..., N, D, H, W, Cin, Cout, dtype)
            d["flops"] += 2.0 * n_ * d_ * h_ * w_ * 27 * ci * co
    out["per"] = per
    del gs, model, opt, resident
    return out


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full) of the dominant kernel's
# largest launch (up_tr64.ops.0 forward at b=32) and its algorithmic bytes: CONSTANTS from the
# committed capture profiles/r02z_ncu_tensor_kernels.md (igemm_roll_kernel, the variant of the family that runs this
# launch since round 2), not re-measured by a bench run
NCU_TRAFFIC = {"bf16": (1.097394e9 + 0.506811e9, 1.640e9), "fp32": (2.195162e9 + 1.040620e9, 3.279e9)}


def roofline_of(r, precision, peaks, peak_src, dev):
    """Per-kernel roofline + breakdown of one measured precision mode."""
    per = r["per"]
    bf16_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    extra = {}
    if precision == "bf16":
        peak = bf16_peak
        src = peak_src + ", sustained bf16 (cuBLAS)"
    else:
        # kind::tf32 runs on the same tcgen05 datapath at K = 8 instead of 16 per instruction: its
        # hardware rate is half the bf16 rate.  MEASURED_PEAKS.json has no TF32 entry, so the
        # denominator is DERIVED; cuBLAS's TF32 GEMM is measured live beside it (it sits below the
        # derived figure on this pool) and the larger of the two is used.
        cublas_tf32 = measure_tf32_peak(dev)
        peak = max(0.5 * bf16_peak, cublas_tf32)
        src = ("DERIVED: max(half of the sustained bf16 peak [" + peak_src + "], cuBLAS TF32 8192^3 GEMM "
               "measured live in this run for 1 s)")
        extra = {"cublas_tf32_tflops_live": cublas_tf32, "half_bf16_sustained_tflops": 0.5 * bf16_peak}
    traffic, algo_bytes = NCU_TRAFFIC[precision]
    if VOL != (64, 64, 32):
        traffic = None                      # the committed ncu capture is of the 64x64x32 workload
    fam = ("pcrl_conv3d_k3_fprop", "pcrl_conv3d_k3_dgrad", "pcrl_conv3d_k3_dgrad_unshuffled")
    kmajor_ms = sum(per[k]["ms"] for k in fam if k in per)
    kmajor_fl = sum(per[k]["flops"] for k in fam if k in per)
    n_kmajor = sum(per[k]["n"] for k in fam if k in per)
    achieved = kmajor_fl / (kmajor_ms / 1e3) / 1e12 if kmajor_ms > 0 else 0.0
    w = per.get("pcrl_conv3d_k3_wgrad", {"ms": 0.0, "flops": 0.0, "n": 0})
    conv_ms, conv_fl = kmajor_ms + w["ms"], kmajor_fl + w["flops"]
    step_ms_prof = sum(d["ms"] for d in per.values())
    breakdown = {k.replace("pcrl_", ""): {"ms": round(v["ms"], 3), "n": v["n"],
                                          **({"tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1),
                                              "frac": round(v["flops"] / (v["ms"] / 1e3) / 1e12 / peak, 3)} if v["flops"] else {})}
                 for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])}
    fl = flops_per_sample()
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic,
            "traffic_note": ("CONSTANT from profiles/r02z_ncu_tensor_kernels.md (ncu --set full, dram__bytes_read.sum + "
                             "dram__bytes_write.sum of this kernel's largest launch, up_tr64.ops.0 forward at b=32; "
                             "algorithmic bytes %.3f GB); not re-measured by this run" % (algo_bytes / 1e9)),
            **extra,
            "kernel": "igemm_kmajor_kernel (3x3x3 conv forward + data gradient)",
            "launches": n_kmajor, "kernel_ms_per_step": kmajor_ms,
            "timing": "CUDA events around every launch of one serialised eager step (side-stream overlap off)",
            "share_of_step": kmajor_ms / step_ms_prof if step_ms_prof else None,
            "wgrad_kernel": {"kernel": "igemm_mnmajor_kernel (3x3x3 weight gradient)", "launches": w["n"],
                             "achieved": w["flops"] / (w["ms"] / 1e3) / 1e12 if w["ms"] else None,
                             "frac": w["flops"] / (w["ms"] / 1e3) / 1e12 / peak if w["ms"] and peak else None,
                             "ms_per_step": w["ms"]},
            "all_conv3_kernels": {"achieved": conv_fl / (conv_ms / 1e3) / 1e12 if conv_ms else None,
                                  "frac": conv_fl / (conv_ms / 1e3) / 1e12 / peak if conv_ms and peak else None,
                                  "ms_per_step": conv_ms},
            "whole_step": {"tflops": r["value"] / r.get("world", 1) * fl / 1e12,
                           "frac_of_this_peak": r["value"] / r.get("world", 1) * fl / 1e12 / peak if peak else None,
                           "frac_of_bf16_burst_peak": (r["value"] / r.get("world", 1) * fl / 1e12 / peaks["bf16_tflops"]
                                                       if peaks.get("bf16_tflops") else None)},
            "peak_source": src}
    return roof, breakdown


def run_ours(args):
    import torch.distributed as dist
    from pcrlv2_b200 import train_3d as T

    rank, world, dev = T.init_distributed()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N>1")
    B = args.batch
    host = [make_batch(B, 1000 * rank + i, True) for i in range(2)]
    main_p = args.precision
    other_p = "bf16" if main_p == "fp32" else "fp32"
    import gc
    r = measure(args, main_p, host, rank, world, dev)
    r["world"] = world
    gc.collect()
    torch.cuda.empty_cache()
    also = None
    if not args.no_also:
        also = measure(args, other_p, host, rank, world, dev)
        also["world"] = world
        gc.collect()
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    roof, breakdown = roofline_of(r, main_p, peaks, peak_src, dev)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, cms, sample = cpu_oracle_rate(2, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    fl = flops_per_sample()
    value = r["value"]
    config = {"workload": WORKLOADS[main_p].format(B=B),
              "global_batch": world * B, "parallelism": f"dp{world}",
              "l2": "activation working set per step is tens of GB >> 126 MB L2; two input batches alternate",
              "algorithmic_gflop_per_sample": round(fl / 1e9, 2), "step": r["step"]}
    line = {
        "metric": METRIC if args.workload == "luna64" else "LUNA 128x128x64 pretrain volumes/sec (configs[3])",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if main_p == "bf16" else "tf32",
        "data": "synthetic",
        "config": config,
        "overall_tflops": value * fl / 1e12,
        "roofline": roof,
        "kernel_breakdown_ms": breakdown,
        "cpu_baseline": cpu,
        "e2e": r["e2e"],
        "gpu_launches": r["gpu_launches"],
        "launches_per_step": r["launches_per_step"],
        "host_ms_per_step": r["host_ms_per_step"],
        "eager_step": r.get("eager"),
        "clocks": r["clocks"],
    }
    if world > 1:
        line["replicas_in_sync"] = r.get("replicas_in_sync")
        config["replicas_in_sync"] = r.get("replicas_in_sync")
    if also is not None:
        roof2, breakdown2 = roofline_of(also, other_p, peaks, peak_src, dev)
        second = {"dtype": "bf16" if other_p == "bf16" else "tf32",
                  "workload": WORKLOADS[other_p].format(B=B), "value": also["value"], "unit": UNIT,
                  "ms_per_step": also["ms_per_step"], "host_ms_per_step": also["host_ms_per_step"],
                  "launches_per_step": also["launches_per_step"], "e2e": also["e2e"],
                  "eager_step": also.get("eager"), "overall_tflops": also["value"] * fl / 1e12,
                  "replicas_in_sync": also.get("replicas_in_sync"), "clocks": also["clocks"]}
        line["also"] = {**second, "roofline": roof2, "kernel_breakdown_ms": breakdown2}
        # the same numbers inside keys the driver's record keeps (it drops unknown top-level keys):
        # the per-GPU shard of configs[2] and its roofline
        key = "configs2_bf16_shard" if other_p == "bf16" else "configs1_fp32"
        config[key] = {k: second[k] for k in ("value", "unit", "ms_per_step", "host_ms_per_step", "overall_tflops",
                                              "replicas_in_sync")}
        config[key]["e2e_value"] = also["e2e"]["value"]
        roof[other_p] = {k: roof2[k] for k in ("achieved", "peak", "frac", "unit", "kernel_ms_per_step", "share_of_step",
                                               "wgrad_kernel", "all_conv3_kernels", "whole_step", "peak_source")}
        roof[other_p]["kernel_breakdown_ms"] = breakdown2
    roof["kernel_breakdown_ms"] = breakdown
    print_line(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (default 32; 8 for --workload large)")
    ap.add_argument("--workload", default="luna64", choices=["luna64", "large", "chest2d"],
                    help="luna64: 64x64x32 crops (configs[1]/[2], the headline); large: configs[3], 128x128x64 "
                         "crops, b=8, bf16 -- its line is recorded under profiles/, it is not the headline; "
                         "chest2d: configs[4], the 2-D path (b=8 per GPU, bf16 unless --precision is given; "
                         "tools/bench_2d.py does the measuring)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["bf16", "fp32"],
                    help="activation storage / tensor-core operand type of the headline measurement "
                         "(fp32 = fp32 storage + TF32 MMAs: configs[1]; bf16: configs[2] shard)")
    ap.add_argument("--eager", action="store_true",
                    help="time the eager step (Python launches) instead of the captured CUDA graph")
    ap.add_argument("--no-also", action="store_true",
                    help="skip the device-resident measurement at the other precision ('also' key)")
    args = ap.parse_args()
    if args.workload == "chest2d":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_2d
        argv = ["--steps", str(args.steps), "--warmup", str(args.warmup), "--batch", str(args.batch or 8), "--skip_torch",
                "--precision", args.precision if "--precision" in sys.argv else "bf16"] + (["--eager"] if args.eager else [])
        return bench_2d.main(argv)
    set_workload(args.workload)
    if args.batch is None:
        args.batch = 8 if args.workload == "large" else 32
    if args.workload == "large" and "--precision" not in sys.argv:
        args.precision = "bf16"
    # stdout carries exactly ONE line (the JSON): anything libraries print there while the job
    # runs (e.g. NCCL's version banner) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = []
    global print_line
    print_line = buf.append
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            if not torch.cuda.is_available():
                raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in buf:
        print(line, flush=True)


if __name__ == "__main__":
    main()
