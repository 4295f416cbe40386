"""Sweep of the weight-gradient kernel's K-stage shape (merged rows per stage x pipeline depth) at the
bench layer shapes.  PCRL_WGRAD_NROWS / PCRL_WGRAD_STAGES are read at every launch, so one process
covers the grid.  Prints ms per (layer, precision, stages, nrows); the launcher lowers nrows until the
stages fit shared memory, PCRL_WGRAD_VERBOSE=1 shows what it chose."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcrlv2_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
LAYERS = [("down_tr64.ops.1", 32, 64, 1), ("down_tr128.ops.0", 64, 64, 2), ("down_tr128.ops.1", 64, 128, 2),
          ("down_tr256.ops.0", 128, 128, 4), ("down_tr256.ops.1", 128, 256, 4), ("down_tr512.ops.0", 256, 256, 8),
          ("down_tr512.ops.1", 256, 512, 8), ("up_tr256.ops.0", 512, 256, 4), ("up_tr256.ops.1", 256, 256, 4),
          ("up_tr128.ops.0", 256, 128, 2), ("up_tr128.ops.1", 128, 128, 2), ("up_tr64.ops.0", 128, 64, 1),
          ("up_tr64.ops.1", 64, 64, 1)]
NROWS = {33: [2, 3, 4, 5, 6, 7, 8], 17: [4, 6, 8, 10, 12, 14, 15, 16], 9: [8, 12, 16, 20, 24, 28, 30], 5: [13, 20, 26, 32, 40, 52]}


def timeit(fn, reps=6):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for DT in (torch.float32, torch.bfloat16):
    for name, cin, cout, s in LAYERS:
        d, h, w = 64 // s, 64 // s, 32 // s
        x = torch.randn(B, d, h + 1, w, cin, device="cuda").to(DT); x[:, :, 0] = 0
        dy = torch.randn(B, d, h + 1, w, cout, device="cuda").to(DT); dy[:, :, 0] = 0
        gpk = torch.zeros(27, cout, cin, device="cuda")
        fl = 2.0 * B * d * h * w * 27 * cin * cout
        os.environ.pop("PCRL_WGRAD_NROWS", None); os.environ.pop("PCRL_WGRAD_STAGES", None)
        t0 = timeit(lambda: K.conv3d_k3_wgrad(dy, x, out=gpk))
        res = []
        for st in (2, 3):
            for nr in NROWS[w + 1]:
                os.environ["PCRL_WGRAD_NROWS"] = str(nr); os.environ["PCRL_WGRAD_STAGES"] = str(st)
                try:
                    t = timeit(lambda: K.conv3d_k3_wgrad(dy, x, out=gpk))
                except Exception as e:
                    t = float("nan")
                res.append((t, st, nr))
        best = min(r for r in res if r[0] == r[0])
        print(f"{'fp32' if DT == torch.float32 else 'bf16'} {name:18s} default {t0:.3f} ms {fl/t0/1e9:6.0f} TF | best {best[0]:.3f} ms "
              f"{fl/best[0]/1e9:6.0f} TF (stages {best[1]}, nrows {best[2]}) | "
              + " ".join(f"s{st}n{nr}:{t:.3f}" for t, st, nr in res), flush=True)
