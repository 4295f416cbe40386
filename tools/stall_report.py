"""Stall accounting of igemm_kmajor_kernel per layer (needs the -DPCRL_TIMING build:
PCRL_EXTRA_FLAGS=-DPCRL_TIMING PCRL_OUT=../libpcrl_b200_timing.so PCRL_BUILD_DIR=build_timing
pcrlv2_b200/csrc/build.sh; run with PCRL_B200_LIB=pcrlv2_b200/libpcrl_b200_timing.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pcrlv2_b200 import kernels as K, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
only = sys.argv[2] if len(sys.argv) > 2 else ""
LAYERS = [("down_tr64.ops.1", 32, 64, 1), ("down_tr128.ops.1", 64, 128, 2), ("down_tr256.ops.1", 128, 256, 4),
          ("up_tr256.ops.0", 512, 256, 4), ("up_tr128.ops.0", 256, 128, 2), ("up_tr128.ops.1", 128, 128, 2),
          ("up_tr64.ops.0", 128, 64, 1), ("up_tr64.ops.1", 64, 64, 1)]
lib = _lib.lib()
lib.pcrl_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
NB = 148
names = ["mma_loop", "w_a_full", "w_b_full", "w_acc_empty", "slabprod_w_empty", "filtprod_w_empty", "epi_w_acc_full", "epi_loop"]
for name, cin, cout, s in LAYERS:
    if only and only not in name: continue
    d, h, w = 64 // s, 64 // s, 32 // s
    x = torch.randn(B, d, h + 1, w, cin, device="cuda").to(torch.bfloat16); x[:, :, 0] = 0
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.02
    wf, wd = K.pack_conv3_weights(wt)
    stats = torch.zeros(1, cout, 2, dtype=torch.float64, device="cuda")
    for _ in range(2):
        K.conv3d_k3_fprop(x, wf, stats)
    torch.cuda.synchronize()
    buf = np.zeros((NB, 8), dtype=np.uint64)
    rc = lib.pcrl_debug_timing(buf.ctypes.data, NB)
    assert rc == 0
    m = buf.astype(np.float64).mean(0)
    fl = 2.0 * B * d * h * w * 27 * cin * cout
    ideal = fl / NB / 8192.0
    print(f"{name:18s} Cin {cin:3d} Cout {cout:3d}: loop {m[0]/1e3:8.0f} kclk (ideal MMA {ideal/1e3:7.0f} = {100*ideal/m[0]:4.1f}%) | MMA warp waits: a_full {100*m[1]/m[0]:4.1f}%  b_full {100*m[2]/m[0]:4.1f}%  acc_empty {100*m[3]/m[0]:4.1f}% | producers idle: slab {100*m[4]/m[0]:4.1f}% filt {100*m[5]/m[0]:4.1f}% | epi idle {100*m[6]/max(m[7],1):4.1f}%", flush=True)
