"""Python wrappers of the 2-D path's entry points (csrc/planar.cu + the plain tensor-core GEMMs).

Activations: H-padded NHWC ``[N, 1, H+1, W, C]`` (the 3-D layout with D = 1) in the storage type
(bf16 or fp32/TF32); the 3-channel network input, masks and targets are plain fp32 NCHW.
Convolution = im2col -> GEMM (+ BatchNorm statistics in the epilogue) -> col2im for the data gradient.
"""
from __future__ import annotations

import torch

from . import _lib
from .kernels import BF16, F32X, _code, _dt, _chk, gemm_nt, gemm_tn, split3


def _round_up(a, b):
    return (a + b - 1) // b * b


def out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def dims2(p):
    """(N, H, W, C) of an H-padded NHWC activation."""
    n, d, h1, w, c = p.shape
    assert d == 1
    return n, h1 - 1, w, c


def round_operand(w32, dtype):
    """fp32 weights -> GEMM operand in the storage type (fp32: cvt.rna.tf32 of the magnitude bits)."""
    if dtype == torch.float32:
        return ((w32.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return w32.to(dtype).contiguous()


def pack_conv2d_weights(w, cs, dtype, exact=False):
    """nn.Conv2d.weight (Cout, Cin, k, k) fp32 -> (wmat [CoutP][Kp], wt [Kp][CoutP]) GEMM operands.
    K index = (ky*k + kx)*cs + c with ``cs`` the channel stride of the input activation (>= Cin; the channels
    above Cin and the K padding are zero); CoutP = Cout rounded up to a multiple of 32."""
    cout, cin, k, _ = w.shape
    coutp = _round_up(cout, 32)
    kp = _round_up(k * k * cs, 64)
    wmat = torch.empty((coutp, kp), dtype=dtype, device=w.device)
    wt = torch.empty((kp, coutp), dtype=dtype, device=w.device)
    _lib.call("pcrl_pack_conv2d_weights", w.detach().contiguous(), wmat, wt, cout, cin, k, cs, coutp, kp,
              _code(dtype, exact))
    if exact:       # 3xTF32: weight side of the split, (hi, hi, lo) along K
        return split3(wmat, 1), split3(wt, 1)
    return wmat, wt


def unpack_conv2d_wgrad(dwmat, cout, cin, k, cs, transposed=False):
    """[CoutP][Kp] fp32 (``transposed``: [Kp][CoutP]) -> (Cout, Cin, k, k)."""
    kp, coutp = (dwmat.shape if transposed else dwmat.shape[::-1])
    g = torch.empty((cout, cin, k, k), dtype=torch.float32, device=dwmat.device)
    _lib.call("pcrl_unpack_conv2d_wgrad", dwmat, g, cout, cin, k, cs, coutp, kp, int(transposed))
    return g


def im2col2d(x, k, s, p, dtype, image=False, cs=None, exact=False):
    """x: H-padded NHWC activation, or (image=True) the fp32 NCHW network input.  Returns (col, Ho, Wo)."""
    if image:
        _chk(x, torch.float32)
        n, c, h, w = x.shape
    else:
        _chk(x, dtype)
        n, h, w, c = dims2(x)
    ho, wo = out_size(h, k, s, p), out_size(w, k, s, p)
    kp = _round_up(k * k * c, 64)
    col = torch.empty((n * (ho + 1) * wo, kp), dtype=dtype, device=x.device)
    _lib.call("pcrl_im2col2d", x, col, n, h, w, c, k, s, p, ho, wo, kp, int(image), _code(dtype, exact))
    return col, ho, wo


def col2im2d(dcol, n, h, w, c, k, s, p, exact=False):
    ho, wo = out_size(h, k, s, p), out_size(w, k, s, p)
    dx = torch.empty((n, 1, h + 1, w, c), dtype=dcol.dtype, device=dcol.device)
    _lib.call("pcrl_col2im2d", dcol, dx, n, h, w, c, k, s, p, ho, wo, dcol.shape[1], _dt(dcol, exact))
    return dx


def gemm_nt_stats(a, b, stats, exact=False):
    """C = A * B^T in the storage type with per-column (sum, sum of squares) ADDED to ``stats`` [1][cols][2] fp64.
    ``exact``: ``b`` is already split (pack_conv2d_weights), ``a`` is split here: 3xTF32, fp32-equivalent products."""
    _chk(a), _chk(b, a.dtype)
    if exact:
        a = split3(a, 0)
    rows, k = a.shape
    cols = b.shape[0]
    c = torch.empty((rows, cols), dtype=a.dtype, device=a.device)
    _lib.call("pcrl_gemm_nt_stats", a, b, c, stats, rows, k, cols, _dt(a, exact))
    return c


def gemm_nt_any(a, b, exact=False):
    """C = A * B^T, result in the storage type (fp32 for fp32 storage); ``exact`` as in gemm_nt_stats."""
    if not exact:
        return gemm_nt(a, b, out_fp32=False)
    a = split3(a, 0)
    rows, k = a.shape
    cols = b.shape[0]
    c = torch.empty((rows, cols), dtype=torch.float32, device=a.device)
    _lib.call("pcrl_gemm_nt", a, b, c, None, rows, k, cols, cols, 1, F32X)
    return c


def _gemm_tn_any(a, b, exact):
    if not exact:
        return gemm_tn(a, b)
    rows, p = a.shape
    out = torch.zeros((p, b.shape[1]), dtype=torch.float32, device=a.device)
    # the reduction runs over the rows: the three partial products are stacked along them
    _lib.call("pcrl_gemm_tn", split3(a, 1, stack=True), split3(b, 0, stack=True), out, 3 * rows, p, b.shape[1], F32X)
    return out


def conv2d_wgrad(dy2d, col, cout, cin, k, cs, exact=False):
    """dW (Cout, Cin, k, k) fp32 = dY^T * col on the tensor cores (operand roles swapped for CoutP = 32, where the
    M = 64 minimum of the MMA would be half empty), unpacked from the GEMM layout."""
    coutp = dy2d.shape[1]
    if coutp % 64 == 0:
        return unpack_conv2d_wgrad(_gemm_tn_any(dy2d, col, exact), cout, cin, k, cs)
    return unpack_conv2d_wgrad(_gemm_tn_any(col, dy2d, exact), cout, cin, k, cs, transposed=True)


def maxpool_fwd(x, exact=False):
    n, h, w, c = dims2(x)
    y = torch.empty((n, 1, (h - 1) // 2 + 2, (w - 1) // 2 + 1, c), dtype=x.dtype, device=x.device)
    _lib.call("pcrl_maxpool2d_3x3s2_fwd", x, y, n, h, w, c, _dt(x, exact))
    return y


def maxpool_bwd(x, dy, exact=False):
    n, h, w, c = dims2(x)
    dx = torch.empty_like(x)
    _lib.call("pcrl_maxpool2d_3x3s2_bwd", x, dy, dx, n, h, w, c, _dt(x, exact))
    return dx


def add_relu(a, b, op=0, exact=False):
    _chk(a), _chk(b, a.dtype)
    out = torch.empty_like(a)
    _lib.call("pcrl_add_relu", a, b, out, a.numel(), op, _dt(a, exact))
    return out


def up_nearest_fwd(x, exact=False):
    n, h, w, c = dims2(x)
    y = torch.empty((n, 1, 2 * h + 1, 2 * w, c), dtype=x.dtype, device=x.device)
    _lib.call("pcrl_upsample_nearest2x_fwd", x, y, n, h, w, c, _dt(x, exact))
    return y


def up_nearest_bwd(g, exact=False):
    n, h2, w2, c = dims2(g)
    dx = torch.empty((n, 1, h2 // 2 + 1, w2 // 2, c), dtype=g.dtype, device=g.device)
    _lib.call("pcrl_upsample_nearest2x_bwd", g, dx, n, h2 // 2, w2 // 2, c, _dt(g, exact))
    return dx


def bilinear_fwd(x, sf):
    _chk(x, torch.float32)
    n, c, h, w = x.shape
    y = torch.empty((n, c, h * sf, w * sf), dtype=torch.float32, device=x.device)
    _lib.call("pcrl_bilinear2d_fwd", x, y, n * c, h, w, sf)
    return y


def bilinear_bwd(g, sf):
    _chk(g, torch.float32)
    n, c, h2, w2 = g.shape
    dx = torch.zeros((n, c, h2 // sf, w2 // sf), dtype=torch.float32, device=g.device)
    _lib.call("pcrl_bilinear2d_bwd", g, dx, n * c, h2 // sf, w2 // sf, sf)
    return dx


def conv_c3_fwd(a, w, bias, c, exact=False):
    """Conv2d(c -> 3, k, padding k//2) + bias on an H-padded activation (channels 0..c-1) -> fp32 NCHW."""
    _chk(a), _chk(w, torch.float32)
    n, h, wd, cs = dims2(a)
    k = w.shape[-1]
    out = torch.empty((n, 3, h, wd), dtype=torch.float32, device=a.device)
    _lib.call("pcrl_conv2d_c3_fwd", a, w, bias, out, n, h, wd, c, cs, k, _dt(a, exact))
    return out


def conv_c3_bwd(a, w, dout, c, need_da=True, exact=False):
    _chk(a), _chk(dout, torch.float32)
    n, h, wd, cs = dims2(a)
    k = w.shape[-1]
    da = torch.empty_like(a) if need_da else None
    dw = torch.zeros_like(w)
    db = torch.zeros((3,), dtype=torch.float32, device=a.device)
    _lib.call("pcrl_conv2d_c3_bwd", a, w, dout, da, dw, db, n, h, wd, c, cs, k, _dt(a, exact))
    return da, dw, db


__all__ = [n for n in dir() if not n.startswith("_")]
