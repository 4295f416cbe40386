"""ConvTranspose3d(k=2,s=2) forward and backward at the three decoder shapes of the bench workload (b=32):
ms, TFLOP/s and the HBM rate the algorithmic bytes imply (read x + write the 8x larger output)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcrlv2_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
only = sys.argv[2] if len(sys.argv) > 2 else ""
DT = torch.float32 if os.environ.get("PCRL_PREC", "bf16") == "fp32" else torch.bfloat16
SHAPES = [("up_tr256.up_conv", 512, 512, 8), ("up_tr128.up_conv", 256, 256, 4), ("up_tr64.up_conv", 128, 128, 2)]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, cin, cout, s in SHAPES:
    if only and only not in name:
        continue
    d, h, w = 64 // s, 64 // s, 32 // s          # coarse dims
    x = torch.randn(B, d, h + 1, w, cin, device="cuda").to(DT); x[:, :, 0] = 0
    wt = torch.randn(cin, cout, 2, 2, 2, device="cuda") * 0.05
    wf, wd = K.pack_convT_weights(wt, dtype=DT)
    bias = torch.zeros(cout, device="cuda")
    g = torch.randn(B, 2 * d, 2 * h + 1, 2 * w, cout, device="cuda").to(DT); g[:, :, 0] = 0
    elt = 4 if DT == torch.float32 else 2
    fl = 2.0 * B * d * h * w * 8 * cin * cout
    by = (B * d * h * w * cin + B * 8 * d * h * w * cout) * elt
    t1 = timeit(lambda: K.convT_fprop(x, wf, bias))
    t2 = timeit(lambda: K.convT_bwd(g, x, wd))
    print(f"{name:18s} {cin}->{cout} coarse {d}x{h}x{w}: fprop {t1:6.3f} ms {fl/t1/1e9:6.1f} TF {by/t1/1e6:6.0f} GB/s | "
          f"bwd (dx + dw + db) {t2:6.3f} ms {2*fl/t2/1e9:6.1f} TF", flush=True)
