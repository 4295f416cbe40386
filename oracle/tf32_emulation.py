"""TF32-operand emulation of the oracle.  TEST INFRASTRUCTURE ONLY (see pcrlv2_oracle.py).

precision='fp32' of the CUDA path keeps every tensor in fp32 in HBM and feeds the tensor cores
kind::tf32 operands: activations, weights and activation gradients are rounded to a 10-bit mantissa
(round to nearest, ties away -- cvt.rna.tf32.f32) where they are stored as operands of a
convolution, products accumulate in fp32.  That is also what the reference itself computes on an
Ampere-or-newer GPU (torch's default torch.backends.cudnn.allow_tf32 = True).  This module restates
the reference forward (models/pcrlv2_model_3d.py:112-133) with that rounding inserted at the
operands of every Conv3d / ConvTranspose3d (forward: x and w; backward: dy), so tests can separate

   CUDA  vs  emulation   : kernel correctness
   emulation vs fp32     : the inherent cost of TF32 operands in the reference's own arithmetic
                           (BatchNorm-backward cancellation amplifies it towards the first layers)
"""
import torch
import torch.nn.functional as F

from . import pcrlv2_oracle as orc


def rna_tf32(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest value with a 10-bit mantissa, ties away from zero (sign-magnitude add)."""
    i = x.contiguous().view(torch.int32)
    out = ((i + 0x1000) & ~0x1FFF).view(torch.float32)          # magnitude bits are the low 31: the
    return torch.where(torch.isfinite(x), out, x)               # add rounds the magnitude for either sign


class _OperandConv(torch.autograd.Function):
    """y = op(rna(x), rna(w)); backward uses rna(dy) for both gradients (what the dgrad / wgrad
    kernels read)."""

    @staticmethod
    def forward(ctx, x, w, kind):
        xr, wr = rna_tf32(x), rna_tf32(w)
        ctx.save_for_backward(xr, wr)
        ctx.kind = kind
        return F.conv3d(xr, wr, None, padding=1) if kind == "conv" else F.conv_transpose3d(xr, wr, None, stride=2)

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        gr = rna_tf32(g)
        with torch.enable_grad():
            xa, wa = xr.detach().requires_grad_(True), wr.detach().requires_grad_(True)
            y = F.conv3d(xa, wa, None, padding=1) if ctx.kind == "conv" else F.conv_transpose3d(xa, wa, None, stride=2)
            dx, dw = torch.autograd.grad(y, (xa, wa), gr)
        return dx, dw, None


def _conv(x, w):
    return _OperandConv.apply(x, w, "conv")


def _norm(y, w, b, norm):
    if norm == "bn":
        return F.batch_norm(y, None, None, w, b, True, 0.1, 1e-5)
    return F.instance_norm(y, None, None, w, b, True, 0.1, 1e-5)


def _luconv(x, sd, prefix, norm):
    y = _conv(x, sd[f"{prefix}.conv1.weight"])        # the bias cancels in the norm
    return F.relu(_norm(y, sd[f"{prefix}.bn1.weight"], sd[f"{prefix}.bn1.bias"], norm))


def forward(sd, x, local=False, norm="bn"):
    """Train-mode forward with TF32 operand rounding; BN buffers are not updated."""
    h = x                                             # the Cin=1 stem is a CUDA-core fp32 kernel
    for i, name in enumerate(orc.DOWN):
        if i > 0:
            h = F.max_pool3d(h, 2)
        if i == 0:
            y = F.conv3d(h, sd[f"{name}.ops.0.conv1.weight"], None, padding=1)
            h = F.relu(_norm(y, sd[f"{name}.ops.0.bn1.weight"], sd[f"{name}.ops.0.bn1.bias"], norm))
        else:
            h = _luconv(h, sd, f"{name}.ops.0", norm)
        h = _luconv(h, sd, f"{name}.ops.1", norm)
    feats, masks = [], []
    for name in orc.UP:
        h = _OperandConv.apply(h, sd[f"{name}.up_conv.weight"], "convT") + sd[f"{name}.up_conv.bias"].view(1, -1, 1, 1, 1)
        h = _luconv(h, sd, f"{name}.ops.0", norm)
        h = _luconv(h, sd, f"{name}.ops.1", norm)
        pro = F.adaptive_avg_pool3d(h, (1, 1, 1)).view(h.shape[0], -1)
        pro = F.batch_norm(pro, None, None, sd[f"{name}.bn.weight"], sd[f"{name}.bn.bias"], True, 0.1, 1e-5)
        t = F.linear(pro, sd[f"{name}.predictor_head.0.weight"], sd[f"{name}.predictor_head.0.bias"])
        t = F.relu(F.batch_norm(t, None, None, sd[f"{name}.predictor_head.1.weight"],
                                sd[f"{name}.predictor_head.1.bias"], True, 0.1, 1e-5))
        pre = F.linear(t, sd[f"{name}.predictor_head.3.weight"], sd[f"{name}.predictor_head.3.bias"])
        ds = f"{name}.deep_supervision_head"
        y1 = _conv(h, sd[f"{ds}.conv1.weight"]) + sd[f"{ds}.conv1.bias"].view(1, -1, 1, 1, 1)
        z = _norm(y1, sd[f"{ds}.bn1.weight"], sd[f"{ds}.bn1.bias"], norm)
        feats.append([pro, pre])
        masks.append(torch.sigmoid(z))
    mm = []
    if not local:
        mm = [F.interpolate(masks[0], scale_factor=4, mode="trilinear"),
              F.interpolate(masks[1], scale_factor=2, mode="trilinear"), masks[2]]
    out = torch.sigmoid(F.conv3d(h, sd["out_tr.final_conv.weight"], sd["out_tr.final_conv.bias"]))
    return out, feats, mm
