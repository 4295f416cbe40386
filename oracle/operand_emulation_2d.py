"""Operand-rounding emulation of the 2-D oracle.  TEST INFRASTRUCTURE ONLY (see pcrlv2_oracle_2d.py).

The CUDA path stores activations in fp32 rounded to TF32 (precision='fp32', what the reference's cuDNN
convolutions compute on a GPU under torch's default allow_tf32) or in bf16 (--amp) and feeds the tensor
cores those operands; products accumulate in fp32.  ``rounding(mode)`` re-runs the SAME oracle code with that
rounding inserted at every convolution: forward operands x and w, the stored output y, and in the backward
pass the operand dy and the stored dx.  Tests then separate

   CUDA  vs  emulation : kernel correctness
   emulation vs fp32   : what TF32 / bf16 operands cost in the reference's own arithmetic -- this 27-convolution
                         network without skip connections amplifies a 2^-11 operand rounding to 2-3e-2 at the
                         output masks (b=4, 64x64, random initialisation)
"""
import contextlib

import torch
import torch.nn.functional as F

from . import pcrlv2_oracle_2d as orc
from .tf32_emulation import rna_tf32


def _round(x, mode):
    if mode == "tf32":
        return rna_tf32(x.float()).to(x.dtype)
    return x.to(torch.bfloat16).to(x.dtype)


class _OperandConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad, mode, x_requires):
        xr, wr = _round(x, mode), _round(w, mode)
        ctx.save_for_backward(xr, wr)
        ctx.cfg = (stride, pad, mode)
        return _round(F.conv2d(xr, wr, None, stride, pad), mode)

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        stride, pad, mode = ctx.cfg
        gr = _round(g, mode)
        with torch.enable_grad():
            xa, wa = xr.detach().requires_grad_(True), wr.detach().requires_grad_(True)
            y = F.conv2d(xa, wa, None, stride, pad)
            dx, dw = torch.autograd.grad(y, (xa, wa), gr)
        return _round(dx, mode), dw, None, None, None, None


class _RoundSTE(torch.autograd.Function):
    """storage rounding of an activation, gradient passed through (and rounded as stored)"""

    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return _round(x, mode)

    @staticmethod
    def backward(ctx, g):
        return _round(g, ctx.mode), None


@contextlib.contextmanager
def rounding(mode):
    """mode: 'tf32' (precision='fp32') or 'bf16'."""
    assert mode in ("tf32", "bf16")

    def impl(x, w, b, stride, pad, kind):
        if kind == "c3":      # fp32 CUDA-core kernel on the STORED (rounded) activation, fp32 weights
            return F.conv2d(_RoundSTE.apply(x, mode), w, b, stride, pad)
        y = _OperandConv2d.apply(x, w, stride, pad, mode, True)
        return y if b is None else y + b.view(1, -1, 1, 1)

    old = orc.CONV_IMPL[0]
    orc.CONV_IMPL[0] = impl
    try:
        yield
    finally:
        orc.CONV_IMPL[0] = old
