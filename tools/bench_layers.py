"""Per-layer timing of the tensor-core kernels at the benchmark shapes (one GPU).
Prints one line per (layer, pass): ms, algorithmic TFLOP/s."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcrlv2_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
only = sys.argv[2] if len(sys.argv) > 2 else ""
DT = torch.float32 if os.environ.get("PCRL_PREC", "bf16") == "fp32" else torch.bfloat16   # fp32 = TF32 MMAs
LAYERS = [("down_tr64.ops.1", 32, 64, 1), ("down_tr128.ops.0", 64, 64, 2), ("down_tr128.ops.1", 64, 128, 2),
          ("down_tr256.ops.0", 128, 128, 4), ("down_tr256.ops.1", 128, 256, 4), ("down_tr512.ops.0", 256, 256, 8),
          ("down_tr512.ops.1", 256, 512, 8), ("up_tr256.ops.0", 512, 256, 4), ("up_tr256.ops.1", 256, 256, 4),
          ("up_tr128.ops.0", 256, 128, 2), ("up_tr128.ops.1", 128, 128, 2), ("up_tr64.ops.0", 128, 64, 1),
          ("up_tr64.ops.1", 64, 64, 1)]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tot = {"fprop": 0, "dgrad": 0, "wgrad": 0}; totf = 0
for name, cin, cout, s in LAYERS:
    if only and only not in name: continue
    d, h, w = 64 // s, 64 // s, 32 // s
    x = torch.randn(B, d, h + 1, w, cin, device="cuda").to(DT); x[:, :, 0] = 0
    dy = torch.randn(B, d, h + 1, w, cout, device="cuda").to(DT); dy[:, :, 0] = 0
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.02
    wf, wd = K.pack_conv3_weights(wt, dtype=DT)
    stats = torch.zeros(1, cout, 2, dtype=torch.float64, device="cuda")
    gpk = torch.zeros(27, cout, cin, device="cuda")
    fl = 2.0 * B * d * h * w * 27 * cin * cout
    t1 = timeit(lambda: K.conv3d_k3_fprop(x, wf, stats))
    t2 = timeit(lambda: K.conv3d_k3_dgrad(dy, wd))
    t3 = timeit(lambda: K.conv3d_k3_wgrad(dy, x, out=gpk))
    tot["fprop"] += t1; tot["dgrad"] += t2; tot["wgrad"] += t3; totf += fl
    print(f"{name:18s} Cin {cin:3d} Cout {cout:3d} {d}x{h}x{w}: fprop {t1:7.3f} ms {fl/t1/1e9:7.1f} TF | dgrad {t2:7.3f} ms {fl/t2/1e9:7.1f} TF | wgrad {t3:7.3f} ms {fl/t3/1e9:7.1f} TF", flush=True)
print("total ms", {k: round(v, 2) for k, v in tot.items()}, "TF", {k: round(totf / v / 1e9, 1) for k, v in tot.items()})
