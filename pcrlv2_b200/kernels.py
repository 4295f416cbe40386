"""Tensor-level wrappers around the C ABI (pcrlv2_b200/_lib.py).

Everything here allocates outputs with torch (device memory, stream ordering) and forwards raw
pointers to libpcrl_b200.so.  Activations are "H-padded NDHWC" bf16 tensors of physical shape
[N, D, H+1, W, C] (row 0 of every plane zero, voxel row h at index h+1) -- see
csrc/common.cuh.  ``pad_ndhwc`` / ``unpad_ndhwc`` convert from / to the reference's NCDHW view.
"""
from __future__ import annotations

import torch

from . import _lib

ACT = {"relu": 0, "prelu": 1, "elu": 2, "sigmoid": 3, "none": 4, "leakyrelu": 5}
BF16 = torch.bfloat16
F32 = torch.float32
# storage types of activations / tensor-core operands (include/pcrl_b200.h: PCRL_DTYPE_*)
DTYPE_CODE = {torch.bfloat16: 0, torch.float32: 1}
F32X = 2     # PCRL_DTYPE_F32X: fp32 stored without tf32 rounding, tensor-core operands split 3xTF32


def _dt(t: torch.Tensor, exact: bool = False) -> int:
    """``exact`` (precision='fp32x3'): fp32 tensors keep all mantissa bits in memory and every
    tensor-core product is evaluated as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo on operands split by
    ``split3`` -- fp32-equivalent arithmetic on the tf32 tensor cores."""
    try:
        code = DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"activations must be bf16 or fp32, got {t.dtype}")
    if exact:
        assert code == 1, "exact (3xTF32) mode stores fp32"
        return F32X
    return code


def _code(dtype, exact: bool = False) -> int:
    return F32X if exact else DTYPE_CODE[dtype]


def split3(t: torch.Tensor, pattern: int, stack: bool = False) -> torch.Tensor:
    """3xTF32 operand split of a contiguous fp32 tensor [..., C] (csrc/streaming.cu:split3_tf32_kernel).
    pattern 0 (activation side) = (hi, lo, hi), pattern 1 (weight side) = (hi, hi, lo);
    ``stack=False`` -> [..., 3C] (parts along the contraction index), ``stack=True`` -> [3*t.shape[0], ...]
    (parts along the leading / row index, for kernels that reduce over rows)."""
    _chk(t, torch.float32)
    c = t.shape[-1]
    rows = t.numel() // c
    if stack:
        out = torch.empty((3 * t.shape[0],) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
    else:
        out = torch.empty(tuple(t.shape[:-1]) + (3 * c,), dtype=torch.float32, device=t.device)
    _lib.call("pcrl_split3_tf32", t, out, rows, c, int(pattern), int(stack))
    return out


def pad_ndhwc(x: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """(N,C,D,H,W) float -> [N,D,H+1,W,C] (bf16 by default) with the zero pad row."""
    n, c, d, h, w = x.shape
    out = torch.zeros((n, d, h + 1, w, c), dtype=dtype, device=x.device)
    out[:, :, 1:] = x.permute(0, 2, 3, 4, 1)
    return out


def unpad_ndhwc(p: torch.Tensor) -> torch.Tensor:
    """[N,D,H+1,W,C] -> (N,C,D,H,W) fp32 (drops the pad row)."""
    return p[:, :, 1:].permute(0, 4, 1, 2, 3).float()


def dims_of(p: torch.Tensor):
    n, d, h1, w, c = p.shape
    return n, d, h1 - 1, w, c


def _chk(t: torch.Tensor, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


# ------------------------------------------------------------------------------ weights
def pack_conv3_weights(w: torch.Tensor, need_dgrad: bool = True, dtype=BF16, exact=False):
    """(Cout,Cin,3,3,3) fp32 -> (wf [27,Cout,Cin], wd [27,Cin,Cout] or None) in ``dtype``;
    ``exact``: the 3xTF32 weight-side split along the contraction index ([27,Cout,3Cin], [27,Cin,3Cout])."""
    _chk(w, torch.float32)
    cout, cin = w.shape[0], w.shape[1]
    wf = torch.empty((27, cout, cin), dtype=dtype, device=w.device)
    wd = torch.empty((27, cin, cout), dtype=dtype, device=w.device) if need_dgrad else None
    _lib.call("pcrl_pack_conv3_weights", w, wf, wd, cout, cin, _code(dtype, exact))
    if exact:
        return split3(wf, 1), (split3(wd, 1) if need_dgrad else None)
    return wf, wd


def unpack_conv3_wgrad(gpk: torch.Tensor) -> torch.Tensor:
    _chk(gpk, torch.float32)
    _, cout, cin = gpk.shape
    g = torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device=gpk.device)
    _lib.call("pcrl_unpack_conv3_wgrad", gpk, g, cout, cin)
    return g


def pack_convT_weights(w: torch.Tensor, dtype=BF16, exact=False):
    """(Cin,Cout,2,2,2) fp32 -> (wf [8*Cout,Cin], wd [Cin,8*Cout]) in ``dtype`` (``exact``: K tripled)."""
    _chk(w, torch.float32)
    cin, cout = w.shape[0], w.shape[1]
    wf = torch.empty((8 * cout, cin), dtype=dtype, device=w.device)
    wd = torch.empty((cin, 8 * cout), dtype=dtype, device=w.device)
    _lib.call("pcrl_pack_convT_weights", w, wf, wd, cin, cout, _code(dtype, exact))
    if exact:
        return split3(wf, 1), split3(wd, 1)
    return wf, wd


def unpack_convT_wgrad(gpk: torch.Tensor, cin: int, cout: int) -> torch.Tensor:
    _chk(gpk, torch.float32)
    g = torch.empty((cin, cout, 2, 2, 2), dtype=torch.float32, device=gpk.device)
    _lib.call("pcrl_unpack_convT_wgrad", gpk, g, cin, cout)
    return g


# ------------------------------------------------------------------------------ 3x3x3 conv
def conv3d_k3_fprop(xp, wf, stats=None, per_sample=False, out_fp32=False, exact=False):
    _chk(xp), _chk(wf, xp.dtype)
    if exact:
        xp = split3(xp, 0)          # [N,D,H+1,W,3Cin] against wf [27,Cout,3Cin]
    n, d, h, w, cin = dims_of(xp)
    cout = wf.shape[1]
    assert wf.shape == (27, cout, cin)
    y = torch.empty((n, d, h + 1, w, cout), dtype=torch.float32 if out_fp32 else xp.dtype,
                    device=xp.device)
    if stats is not None:
        _chk(stats, torch.float64)
    _lib.call("pcrl_conv3d_k3_fprop", xp, wf, y, stats, int(per_sample), int(out_fp32), n, d, h, w,
              cin, cout, _dt(xp, exact))
    return y


def conv3d_k3_dgrad(dyp, wd, exact=False):
    _chk(dyp), _chk(wd, dyp.dtype)
    if exact:
        dyp = split3(dyp, 0)
    n, d, h, w, cout = dims_of(dyp)
    cin = wd.shape[1]
    assert wd.shape == (27, cin, cout)
    dx = torch.empty((n, d, h + 1, w, cin), dtype=dyp.dtype, device=dyp.device)
    _lib.call("pcrl_conv3d_k3_dgrad", dyp, wd, dx, n, d, h, w, cin, cout, _dt(dyp, exact))
    return dx


def conv3d_k3_dgrad_unshuffled(dyp, wd, exact=False):
    """Data gradient written coarse-major for the ConvTranspose that produced the conv input.
    Returns (scratch [N*(D/2)*(H/2+1)*(W/2), 8*Cin] bf16, colsum [Cin,2] fp64)."""
    _chk(dyp), _chk(wd, dyp.dtype)
    if exact:
        dyp = split3(dyp, 0)
    n, d, h, w, cout = dims_of(dyp)
    cin = wd.shape[1]
    rows = n * (d // 2) * (h // 2 + 1) * (w // 2)
    scratch = torch.empty((rows, 8 * cin), dtype=dyp.dtype, device=dyp.device)
    colsum = torch.zeros((cin, 2), dtype=torch.float64, device=dyp.device)
    _lib.call("pcrl_conv3d_k3_dgrad_unshuffled", dyp, wd, scratch, colsum, n, d, h, w, cin, cout, _dt(dyp, exact))
    return scratch, colsum


def convT_bwd_from_scratch(scratch, xp, wd, need_dx=True, exact=False):
    """ConvTranspose3d(k2,s2) gradient GEMMs on an already coarse-major output gradient."""
    n, d, h, w, cin = dims_of(xp)
    cout = scratch.shape[1] // 8
    if exact:
        # the two GEMMs of pcrl_convT3d_k2s2_bwd with split operands: dx = scratch * wd^T (parts along
        # K = 8*Cout), dw = scratch^T * x (parts stacked along the rows)
        rows = scratch.shape[0]
        dx = None
        if need_dx:
            dx = torch.empty((n, d, h + 1, w, cin), dtype=xp.dtype, device=xp.device)
            _lib.call("pcrl_gemm_nt", split3(scratch, 0), wd, dx, None, rows, 24 * cout, cin, cin, 1, F32X)
        dw = torch.zeros((8 * cout, cin), dtype=torch.float32, device=xp.device)
        _lib.call("pcrl_gemm_tn", split3(scratch, 1, stack=True), split3(xp.reshape(rows, cin), 0, stack=True),
                  dw, 3 * rows, 8 * cout, cin, F32X)
        return dx, dw
    dx = torch.empty((n, d, h + 1, w, cin), dtype=xp.dtype, device=xp.device) if need_dx else None
    dw = torch.zeros((8 * cout, cin), dtype=torch.float32, device=xp.device)
    _lib.call("pcrl_convT3d_k2s2_bwd", None, xp, wd, scratch, dx, dw, None, n, d, h, w, cin, cout, _dt(xp))
    return dx, dw


def conv3d_k3_wgrad(dyp, xp, out=None, exact=False):
    """Returns / accumulates into the packed gradient [27,Cout,Cin] fp32."""
    _chk(dyp), _chk(xp, dyp.dtype)
    if exact:     # the reduction runs over (sample, voxel): parts stacked along the sample index
        dyp, xp = split3(dyp, 1, stack=True), split3(xp, 0, stack=True)
    n, d, h, w, cout = dims_of(dyp)
    cin = xp.shape[-1]
    if out is None:
        out = torch.zeros((27, cout, cin), dtype=torch.float32, device=xp.device)
    _lib.call("pcrl_conv3d_k3_wgrad", dyp, xp, out, n, d, h, w, cin, cout, _dt(dyp, exact))
    return out


# ------------------------------------------------------------------------------ stem
def stem_conv_fprop(x, w, stats=None, per_sample=False, dtype=BF16, exact=False):
    """x (N,1,D,H,W) fp32, w (32,1,3,3,3) fp32 -> H-padded [N,D,H+1,W,32] in ``dtype``."""
    _chk(x, torch.float32), _chk(w, torch.float32)
    n, _, d, h, wd_ = x.shape
    assert w.shape[0] == 32 and w.shape[1] == 1
    y = torch.empty((n, d, h + 1, wd_, 32), dtype=dtype, device=x.device)
    _lib.call("pcrl_stem_conv_fprop", x, w, y, stats, int(per_sample), n, d, h, wd_, _code(dtype, exact))
    return y


def stem_conv_wgrad(dyp, x):
    _chk(dyp), _chk(x, torch.float32)
    n, _, d, h, w = x.shape
    dw = torch.zeros((32, 1, 3, 3, 3), dtype=torch.float32, device=x.device)
    _lib.call("pcrl_stem_conv_wgrad", dyp, x, dw, n, d, h, w, _dt(dyp))
    return dw


# ------------------------------------------------------------------------------ ConvTranspose
def convT_fprop(xp, wf, bias, exact=False):
    _chk(xp), _chk(wf, xp.dtype)
    if exact:
        xp = split3(xp, 0)
    n, d, h, w, cin = dims_of(xp)
    cout = wf.shape[0] // 8
    y = torch.empty((n, 2 * d, 2 * h + 1, 2 * w, cout), dtype=xp.dtype, device=xp.device)
    _lib.call("pcrl_convT3d_k2s2_fprop", xp, wf, bias, y, n, d, h, w, cin, cout, _dt(xp, exact))
    return y


def convT_bwd(gp, xp, wd, need_dx=True, need_dw=True):
    """gp: gradient wrt the fine output.  Returns (dx padded bf16, dw_packed fp32, dbias fp32)."""
    _chk(gp)
    n, d2, h2, w2, cout = dims_of(gp)
    d, h, w = d2 // 2, h2 // 2, w2 // 2
    cin = wd.shape[0]
    rows = n * d * (h + 1) * w
    scratch = torch.empty((rows, 8 * cout), dtype=gp.dtype, device=gp.device)
    dx = torch.empty((n, d, h + 1, w, cin), dtype=gp.dtype, device=gp.device) if need_dx else None
    dw = torch.zeros((8 * cout, cin), dtype=torch.float32, device=gp.device) if need_dw else None
    db = torch.zeros((cout,), dtype=torch.float32, device=gp.device)
    _lib.call("pcrl_convT3d_k2s2_bwd", gp, xp if need_dw else None, wd, scratch, dx, dw, db,
              n, d, h, w, cin, cout, _dt(gp))
    return dx, dw, db


# ------------------------------------------------------------------------------ norm + act
def norm_finalize(stats, count, gamma, beta, conv_bias=None, running_mean=None, running_var=None,
                  nbt=None, momentum=0.1, eps=1e-5):
    g, c = stats.shape[0], stats.shape[1]
    dev = stats.device
    scale = torch.empty((g, c), dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    _lib.call("pcrl_norm_finalize", stats, float(count), gamma, beta, conv_bias, running_mean,
              running_var, nbt, float(momentum), float(eps), scale, shift, mean, invstd, g, c)
    return scale, shift, mean, invstd


def norm_act_fwd(yp, scale, shift, act="relu", prelu=None, want_full=True, want_pool=False,
                 want_avg=False, per_sample=False, exact=False):
    _chk(yp)
    n, d, h, w, c = dims_of(yp)
    a = torch.empty_like(yp) if want_full else None
    pool = (torch.empty((n, d // 2, h // 2 + 1, w // 2, c), dtype=yp.dtype, device=yp.device)
            if want_pool else None)
    avg = torch.zeros((n, c), dtype=torch.float32, device=yp.device) if want_avg else None
    _lib.call("pcrl_norm_act_fwd", yp, scale, shift, prelu, a, pool, avg, int(per_sample), ACT[act],
              int(want_pool), n, d, h, w, c, _dt(yp, exact))
    return a, pool, avg


def norm_act_bwd(yp, g1, g2, gavg, scale, shift, mean, invstd, gamma, act="relu", prelu=None,
                 pool=False, per_sample=False, batch_stats=True, exact=False):
    """Returns (dy padded, sums [G,C,3] fp64 = (dbeta, dgamma, dprelu partial)).
    ``batch_stats=False`` (BatchNorm in eval mode: mean / invstd are the running statistics, constants
    of the graph): dy = gamma*invstd*dz without the mean(dz) / xhat*mean(dz*xhat) terms -- an infinite
    element count makes those terms vanish in the apply pass; the parameter-gradient sums are unchanged."""
    _chk(yp)
    n, d, h, w, c = dims_of(yp)
    g = scale.shape[0]
    sums = torch.zeros((g, c, 3), dtype=torch.float64, device=yp.device)
    count = float(d * h * w) if per_sample else float(n * d * h * w)
    if not batch_stats:
        count = float("inf")
    dy = torch.empty_like(yp)
    for p in (0, 1):
        _lib.call("pcrl_norm_act_bwd", yp, g1, g2, gavg, scale, shift, mean, invstd, gamma, prelu,
                  sums, dy, count, int(per_sample), ACT[act], int(pool), p, n, d, h, w, c, _dt(yp, exact))
    return dy, sums


# ------------------------------------------------------------------------------ heads
def head_pack_weights(w3, w1=None, dtype=BF16, exact=False):
    """(1,C,3,3,3) [+ (1,C,1,1,1)] fp32 -> (wext [32,C], wextT [C,32]) in ``dtype`` (``exact``: K tripled)."""
    c = w3.shape[1]
    wext = torch.empty((32, c), dtype=dtype, device=w3.device)
    wext_t = torch.empty((c, 32), dtype=dtype, device=w3.device)
    _lib.call("pcrl_head_pack_weights", w3.contiguous(), None if w1 is None else w1.contiguous(),
              wext, wext_t, c, _code(dtype, exact))
    if exact:
        return split3(wext, 1), split3(wext_t, 1)
    return wext, wext_t


def head_fwd(ap, wext, b3, b1=None, stats=None, per_sample=False, exact=False):
    """1-channel head convolutions of an H-padded activation: T = A * wext^T on the tensor cores,
    then the 27-point gather.  Returns (y1 (N,1,D,H,W) fp32, y0 or None)."""
    _chk(ap), _chk(wext, ap.dtype)
    n, d, h, w, c = dims_of(ap)
    rows = n * d * (h + 1) * w
    t_t = torch.empty((32, rows), dtype=torch.float32, device=ap.device)
    if exact:
        _lib.call("pcrl_gemm_nt", split3(ap, 0), wext, t_t, None, rows, 3 * c, 32, rows, 2, F32X)
    else:
        _lib.call("pcrl_gemm_nt", ap, wext, t_t, None, rows, c, 32, rows, 2, _dt(ap))
    y1 = torch.empty((n, 1, d, h, w), dtype=torch.float32, device=ap.device)
    y0 = torch.empty_like(y1) if b1 is not None else None
    _lib.call("pcrl_head_gather", t_t, b3, b1, y1, y0, stats, int(per_sample), n, d, h, w)
    return y1, y0


def head_bwd(ap, dy1, dy0, wext_t, exact=False):
    """Returns (dA H-padded bf16, dwext [C,32] fp32: columns 0..26 = d w3 (tap order), 27 = d w1)."""
    n, d, h, w, c = dims_of(ap)
    rows = n * d * (h + 1) * w
    d_t = torch.empty((rows, 32), dtype=ap.dtype, device=ap.device)
    _lib.call("pcrl_head_scatter", dy1, dy0, d_t, n, d, h, w, _dt(ap, exact))
    da = torch.empty((n, d, h + 1, w, c), dtype=ap.dtype, device=ap.device)
    if exact:
        _lib.call("pcrl_gemm_nt", split3(d_t, 0), wext_t, da, None, rows, 96, c, c, 1, F32X)
        dwext = torch.zeros((c, 32), dtype=torch.float32, device=ap.device)
        _lib.call("pcrl_gemm_tn", split3(ap.reshape(rows, c), 0, stack=True), split3(d_t, 1, stack=True),
                  dwext, 3 * rows, c, 32, F32X)
        return da, dwext
    _lib.call("pcrl_gemm_nt", d_t, wext_t, da, None, rows, 32, c, c, _dt(ap), _dt(ap))
    dwext = torch.zeros((c, 32), dtype=torch.float32, device=ap.device)
    _lib.call("pcrl_gemm_tn", ap, d_t, dwext, rows, c, 32, _dt(ap))
    return da, dwext


def chan1_sigmoid_fwd(y, scale, shift, per_sample):
    """mask = sigmoid(y*scale + shift); y (N,1,D,H,W) fp32; scale/shift [G] (G = N if per_sample)."""
    mask = torch.empty_like(y)
    n = y.shape[0]
    vol = y.numel() // n
    g, v = (n, vol) if per_sample else (1, y.numel())
    _lib.call("pcrl_chan1_sigmoid_fwd", y, scale, shift, mask, int(per_sample), g, v)
    return mask


def chan1_sigmoid_bwd(y, mask, dmask, mean, invstd, gamma, per_sample, batch_stats=True):
    """Returns (dy, sums [G,3] fp64 with (d beta, d gamma, 0) partials); ``batch_stats`` as in
    norm_act_bwd."""
    n = y.shape[0]
    vol = y.numel() // n
    g, v = (n, vol) if per_sample else (1, y.numel())
    sums = torch.zeros((g, 3), dtype=torch.float64, device=y.device)
    dy = torch.empty_like(y)
    for p in (0, 1):
        _lib.call("pcrl_chan1_sigmoid_bwd", y, mask, dmask, mean, invstd, gamma, sums, dy,
                  float(v) if batch_stats else float("inf"), int(per_sample), p, g, v)
    return dy, sums


def im2col27(x, dtype, exact=False):
    """x (N,1,D,H,W) fp32 -> X27 [rows, 32] in ``dtype`` (H-padded row order, X27[u][tap] = x[u + tap])."""
    _chk(x, torch.float32)
    n, _, d, h, w = x.shape
    x27 = torch.empty((n * d * (h + 1) * w, 32), dtype=dtype, device=x.device)
    _lib.call("pcrl_im2col27", x, x27, n, d, h, w, _code(dtype, exact))
    return x27


def stem_pack_weights(w, dtype):
    """(32,1,3,3,3) fp32 -> [32 out, 32 taps] GEMM operand (taps 27..31 zero), rounded to the operand type."""
    w32 = torch.nn.functional.pad(w.detach().reshape(32, 27), (0, 5)).contiguous()
    if dtype == torch.float32:      # cvt.rna.tf32 (round to nearest, ties away) of the magnitude bits
        return ((w32.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return w32.to(dtype)


def stem_conv_fprop_gemm(x27, wst, dims, stats):
    """Conv3d(1 -> 32) as X27 * Wst^T on the tensor cores with the BatchNorm statistics in the epilogue.
    Returns the H-padded activation [N,D,H+1,W,32] (pad rows are zero because X27's are)."""
    n, d, h, w = dims
    y = torch.empty((n, d, h + 1, w, 32), dtype=x27.dtype, device=x27.device)
    _lib.call("pcrl_gemm_nt_stats", x27, wst, y, stats, x27.shape[0], 32, 32, _dt(x27))
    return y


def stem_conv_wgrad_gemm(dyp, x, exact=False, x27=None):
    """Stem weight gradient on the tensor cores: im2col of the 1-channel input to [rows,32]
    (27 taps; ``x27``: the one the forward pass already made), rows paired into 64-wide operands,
    dW = sum of the diagonal 32x32 blocks of gemm_tn(dY, X27)."""
    _chk(dyp)
    rows = dyp.numel() // 32
    if x27 is None:
        x27 = im2col27(x, dyp.dtype, exact)
    out = torch.zeros((64, 64), dtype=torch.float32, device=dyp.device)
    if exact:
        _lib.call("pcrl_gemm_tn", split3(dyp.reshape(rows // 2, 64), 1, stack=True),
                  split3(x27.reshape(rows // 2, 64), 0, stack=True), out, 3 * (rows // 2), 64, 64, F32X)
    else:
        _lib.call("pcrl_gemm_tn", dyp, x27, out, rows // 2, 64, 64, _dt(dyp))
    dw = out[:32, :32] + out[32:, 32:]
    return dw[:, :27].reshape(32, 1, 3, 3, 3).contiguous()


# ------------------------------------------------------------------------------ GEMMs / SGD
def gemm_nt(a, b, bias=None, out_fp32=True):
    _chk(a), _chk(b, a.dtype)
    rows, k = a.shape
    cols = b.shape[0]
    if out_fp32 == "transposed":
        c = torch.empty((cols, rows), dtype=torch.float32, device=a.device)
        _lib.call("pcrl_gemm_nt", a, b, c, bias, rows, k, cols, rows, 2, _dt(a))
        return c
    out_fp32 = bool(out_fp32) or a.dtype == torch.float32
    c = torch.empty((rows, cols), dtype=torch.float32 if out_fp32 else BF16, device=a.device)
    _lib.call("pcrl_gemm_nt", a, b, c, bias, rows, k, cols, cols, int(out_fp32), _dt(a))
    return c


def gemm_tn(a, b, out=None):
    _chk(a), _chk(b, a.dtype)
    rows, p = a.shape
    q = b.shape[1]
    if out is None:
        out = torch.zeros((p, q), dtype=torch.float32, device=a.device)
    _lib.call("pcrl_gemm_tn", a, b, out, rows, p, q, _dt(a))
    return out


def sgd_flat_dev(params, grads, bufs, seg_off, seg_active, seg_first, hyper, guard=None):
    """sgd_flat with [lr, momentum, weight_decay, grad_scale, skip_threshold] in device memory."""
    _lib.call("pcrl_sgd_flat_dev", params, grads, bufs, seg_off, seg_active, seg_first, seg_active.numel(),
              hyper, guard)


def sgd_flat(params, grads, bufs, seg_off, seg_active, seg_first, lr, momentum, weight_decay,
             grad_scale=1.0):
    nseg = seg_active.numel()
    _lib.call("pcrl_sgd_flat", params, grads, bufs, seg_off, seg_active, seg_first, nseg, float(lr),
              float(momentum), float(weight_decay), float(grad_scale))


# ------------------------------------------------------------------------------ heads / losses (fp32)
def _f32c(t):
    assert t.dtype == torch.float32 and t.is_contiguous(), "expected a contiguous fp32 tensor"
    return t


def bn1d_fwd(x, gamma, beta, running_mean, running_var, nbt, relu, training, momentum=0.1, eps=1e-5):
    """BatchNorm1d over the rows of x (B, C) [+ ReLU]; returns (y, save_mean, save_invstd)."""
    b, c = x.shape
    y = torch.empty_like(_f32c(x))
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    invstd = torch.empty(c, dtype=torch.float32, device=x.device)
    _lib.call("pcrl_bn1d_fwd", x, _f32c(gamma), _f32c(beta), running_mean, running_var, nbt, y, mean, invstd,
              b, c, int(relu), int(training), float(momentum), float(eps))
    return y, mean, invstd


def bn1d_bwd(x, y, dy, gamma, mean, invstd, relu, training):
    b, c = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    _lib.call("pcrl_bn1d_bwd", x, y, _f32c(dy), gamma, mean, invstd, dx, dgamma, dbeta, b, c, int(relu), int(training))
    return dx, dgamma, dbeta


def linear_fwd(x, w, bias):
    b, k = x.shape
    j = w.shape[0]
    y = torch.empty((b, j), dtype=torch.float32, device=x.device)
    _lib.call("pcrl_linear_fwd", _f32c(x), _f32c(w), bias, y, b, k, j)
    return y


def linear_bwd(x, w, dy, need_dx=True):
    b, k = x.shape
    j = w.shape[0]
    dx = torch.empty((b, k), dtype=torch.float32, device=x.device) if need_dx else None
    dw = torch.empty((j, k), dtype=torch.float32, device=x.device)
    db = torch.empty((j,), dtype=torch.float32, device=x.device)
    _lib.call("pcrl_linear_bwd", x, w, _f32c(dy), dx, dw, db, b, k, j)
    return dx, dw, db


def cosine_mean_fwd_bwd(x, y, eps=1e-8, coef=1.0, need_dx=True):
    """coef * mean_b cos(x_b, y_b) (0-dim tensor) and its gradient wrt x."""
    b, c = x.shape
    out = torch.zeros((), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if need_dx else None
    _lib.call("pcrl_cosine_mean_fwd_bwd", _f32c(x), _f32c(y), out, dx, b, c, float(eps), float(coef))
    return out, dx


def mse_fwd(p, t, weight=None):
    """mean((p - t)^2) [* weight[0], a device scalar]."""
    out = torch.zeros((), dtype=torch.float32, device=p.device)
    if weight is None:
        _lib.call("pcrl_mse_fwd", _f32c(p), _f32c(t), out, p.numel())
    else:
        _lib.call("pcrl_mse_scaled_fwd", _f32c(p), _f32c(t), _f32c(weight), out, p.numel())
    return out


def mse_bwd(p, t, g, weight=None):
    dp = torch.empty_like(p)
    if weight is None:
        _lib.call("pcrl_mse_bwd", p, t, _f32c(g), dp, p.numel())
    else:
        _lib.call("pcrl_mse_scaled_bwd", p, t, _f32c(g), _f32c(weight), dp, p.numel())
    return dp


def contrastive_fwd_bwd(pre1, pro1, pre2, pro2, pre_l, pro_l, draws, eps=1e-8):
    """The 1 + 2*n_local cos_loss terms of a step and their gradients in one launch.
    pre1/pro1/pre2/pro2: S tensors [B, C_s] each (S = 3 for the 3-D model, 5 for the 2-D one); pre_l/pro_l:
    S tensors [n_local*B, C_s]; draws: int32 device tensor [1 + 2*n_local].  Returns (out [2] = (loss2,
    local_loss), (dpre1[S], dpre2[S], dpre_l[S]))."""
    import ctypes
    b = pre1[0].shape[0]
    S = len(pre1)
    n_local = pre_l[0].shape[0] // b
    assert draws.dtype == torch.int32 and draws.is_cuda and draws.numel() >= 1 + 2 * n_local
    ins = [_f32c(t) for grp in (pre1, pro1, pre2, pro2, pre_l, pro_l) for t in grp]
    grads = [torch.empty_like(t) for grp in (pre1, pre2, pre_l) for t in grp]
    ptrs = (ctypes.c_void_p * (9 * S))(*[t.data_ptr() for t in ins + grads])
    chans = (ctypes.c_int * S)(*[int(t.shape[1]) for t in pre1])
    out = torch.zeros(2, dtype=torch.float32, device=draws.device)
    _lib.call("pcrl_contrastive_fwd_bwd_s", ptrs, chans, S, b, n_local, draws, out, float(eps))
    return out, (grads[0:S], grads[S:2 * S], grads[2 * S:3 * S])


def sigmoid_fwd(x):
    y = torch.empty_like(_f32c(x))
    _lib.call("pcrl_sigmoid_fwd", x, y, x.numel())
    return y


def sigmoid_bwd(y, dy):
    dx = torch.empty_like(y)
    _lib.call("pcrl_sigmoid_bwd", y, _f32c(dy), dx, y.numel())
    return dx


def upsample_trilinear_fwd(x, sf):
    """x (N,1,D,H,W) fp32 -> (N,1,D*sf,H*sf,W*sf)."""
    n, _, d, h, w = x.shape
    y = torch.empty((n, 1, d * sf, h * sf, w * sf), dtype=torch.float32, device=x.device)
    _lib.call("pcrl_upsample_trilinear_fwd", _f32c(x), y, n, d, h, w, int(sf))
    return y


def upsample_trilinear_bwd(dy, sf):
    n, _, od, oh, ow = dy.shape
    d, h, w = od // sf, oh // sf, ow // sf
    dx = torch.zeros((n, 1, d, h, w), dtype=torch.float32, device=dy.device)
    _lib.call("pcrl_upsample_trilinear_bwd", _f32c(dy), dx, n, d, h, w, int(sf))
    return dx
