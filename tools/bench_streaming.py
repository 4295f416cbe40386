"""Timing of the HBM-bound kernels at the bench shapes (b=32): achieved GB/s on ALGORITHMIC bytes.
   python tools/bench_streaming.py [bf16|fp32] [filter]
Used under ncu for the dram__bytes captures of profiles/ (see tools/run_profiles.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcrlv2_b200 import kernels as K
from pcrlv2_b200 import _lib

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
only = sys.argv[2] if len(sys.argv) > 2 else ""
DT = torch.float32 if prec == "fp32" else torch.bfloat16
E = 4 if prec == "fp32" else 2
B = 32


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()                      # L2 flush between timed launches
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, ms, nbytes):
    print(f"{name:58s} {ms:7.3f} ms  {nbytes / 1e9:6.3f} GB algorithmic  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


shapes = [("64x64x32 C=64", 64, 64, 32, 64), ("64x64x32 C=32", 64, 64, 32, 32), ("32x32x16 C=128", 32, 32, 16, 128),
          ("16x16x8 C=256", 16, 16, 8, 256)]
for tag, d, h, w, c in shapes:
    if only and only not in tag and only not in ("fwd", "bwd", "pool"):
        continue
    n_el = B * d * h * w * c            # real voxels x channels
    y = torch.randn(B, d, h + 1, w, c, device="cuda").to(DT); y[:, :, 0] = 0
    g = torch.randn(B, d, h + 1, w, c, device="cuda").to(DT); g[:, :, 0] = 0
    gp = torch.randn(B, d // 2, h // 2 + 1, w // 2, c, device="cuda").to(DT); gp[:, :, 0] = 0
    scale = torch.rand(1, c, device="cuda") + 0.5
    shift = torch.randn(1, c, device="cuda") * 0.1
    mean = torch.randn(1, c, device="cuda") * 0.1
    invstd = torch.rand(1, c, device="cuda") + 0.5
    gamma = torch.ones(c, device="cuda")
    if not only or only in tag or only == "fwd":
        report(f"norm_act_fwd relu {tag}", timeit(lambda: K.norm_act_fwd(y, scale, shift)), 2 * n_el * E)
        report(f"norm_act_fwd relu+pool {tag}", timeit(lambda: K.norm_act_fwd(y, scale, shift, want_full=False, want_pool=True)),
               n_el * E + n_el * E // 8)
    if not only or only in tag or only == "bwd":
        sums = torch.zeros(1, c, 3, dtype=torch.float64, device="cuda")
        dy = torch.empty_like(y)
        cnt = float(B * d * h * w)

        def p(pass_, pool=False, g1=g):
            _lib.call("pcrl_norm_act_bwd", y, g1, None, None, scale, shift, mean, invstd, gamma, None, sums, dy, cnt,
                      0, 0, int(pool), pass_, B, d, h, w, c, K._dt(y))
        report(f"norm_act_bwd pass0 (sums) {tag}", timeit(lambda: p(0)), 2 * n_el * E)
        report(f"norm_act_bwd pass1 (apply) {tag}", timeit(lambda: p(1)), 3 * n_el * E)
        report(f"norm_act_bwd pool pass0 {tag}", timeit(lambda: p(0, True, gp)), n_el * E + n_el * E // 8)
        report(f"norm_act_bwd pool pass1 {tag}", timeit(lambda: p(1, True, gp)), 2 * n_el * E + n_el * E // 8)
if not only or only == "stem":
    x = torch.randn(B, 1, 64, 64, 32, device="cuda")
    wt = torch.randn(32, 1, 3, 3, 3, device="cuda")
    st = torch.zeros(1, 32, 2, dtype=torch.float64, device="cuda")
    report("stem_conv_fprop 64x64x32", timeit(lambda: K.stem_conv_fprop(x, wt, st, dtype=DT)), B * 131072 * (4 + 32 * E))
if not only or only == "head":
    a = torch.randn(B, 64, 65, 32, 64, device="cuda").to(DT); a[:, :, 0] = 0
    w3 = torch.randn(1, 64, 3, 3, 3, device="cuda"); w1 = torch.randn(1, 64, 1, 1, 1, device="cuda")
    wext, _ = K.head_pack_weights(w3, w1, dtype=DT)
    b3 = torch.zeros(1, device="cuda"); b1 = torch.zeros(1, device="cuda")
    st1 = torch.zeros(1, 1, 2, dtype=torch.float64, device="cuda")
    rows = B * 64 * 65 * 32
    report("head_fwd (GEMM + gather) 64x64x32 C=64", timeit(lambda: K.head_fwd(a, wext, b3, b1, st1)),
           rows * 64 * E + 2 * rows * 32 * 4 + 2 * B * 131072 * 4)
