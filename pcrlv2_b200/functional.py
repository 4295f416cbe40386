"""Autograd nodes of the small fp32 pieces of the step -- the projection / prediction heads and the
loss terms -- on the CUDA kernels of csrc/losses.cu (no ATen arithmetic on the hot path):

  batch_norm1d   nn.BatchNorm1d (+ReLU)        reference models/pcrlv2_model_3d.py:54,56-57,67,69
  linear         nn.Linear                     reference models/pcrlv2_model_3d.py:55,58,69
  cosine_mean    nn.CosineSimilarity()(x, y.detach()).mean()      reference train_3d.py:90-91
  mse_loss       nn.MSELoss()                  reference train_3d.py:135,137
  sigmoid        torch.sigmoid                 reference models/pcrlv2_model_3d.py:79,132
  upsample_trilinear  F.interpolate(scale_factor, mode='trilinear')  reference :125-126
"""
import torch

from . import kernels as K


class _BN1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu, training):
        x = x.contiguous()
        track = training and bn.track_running_stats and bn.running_mean is not None
        use_batch = training or bn.running_mean is None
        momentum = 0.1 if bn.momentum is None else bn.momentum
        y, mean, invstd = K.bn1d_fwd(
            x, gamma.detach().contiguous(), beta.detach().contiguous(),
            bn.running_mean if (track or not use_batch) else None,
            bn.running_var if (track or not use_batch) else None,
            bn.num_batches_tracked if track else None, relu, use_batch, momentum, bn.eps)
        ctx.relu, ctx.use_batch = relu, use_batch
        ctx.save_for_backward(x, y, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, invstd = ctx.saved_tensors
        dx, dgamma, dbeta = K.bn1d_bwd(x, y, dy.contiguous(), gamma.detach().contiguous(), mean, invstd,
                                       ctx.relu, ctx.use_batch)
        return dx, dgamma, dbeta, None, None, None


def batch_norm1d(x, bn, relu=False):
    """``bn`` is the nn.BatchNorm1d parameter container (affine, running statistics)."""
    return _BN1dFn.apply(x, bn.weight, bn.bias, bn, relu, bn.training)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        ctx.save_for_backward(x, w)
        return K.linear_fwd(x, w.detach().contiguous(), b.detach().contiguous() if b is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = K.linear_bwd(x, w.detach().contiguous(), dy.contiguous(), need_dx=ctx.needs_input_grad[0])
        return dx, dw, db


def linear(x, lin):
    return _LinearFn.apply(x, lin.weight, lin.bias)


class _CosineMeanFn(torch.autograd.Function):
    """coef * cosine_similarity(x, y, dim=1, eps).mean(); y is a constant (the reference detaches it)."""

    @staticmethod
    def forward(ctx, x, y, eps, coef):
        out, dx = K.cosine_mean_fwd_bwd(x.contiguous(), y.detach().contiguous(), eps, coef,
                                        need_dx=ctx.needs_input_grad[0])
        ctx.save_for_backward(dx)
        return out

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return (dx * g if dx is not None else None), None, None, None


def cosine_mean(x, y, eps=1e-8, coef=1.0):
    return _CosineMeanFn.apply(x, y, eps, coef)


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t):
        p, t = p.contiguous(), t.contiguous()
        ctx.save_for_backward(p, t)
        return K.mse_fwd(p, t)

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        return K.mse_bwd(p, t, g.contiguous().float()), None


def mse_loss(pred, target):
    if target.requires_grad:
        raise NotImplementedError("mse_loss: the target is a constant on this path")
    return _MSEFn.apply(pred, target.to(pred.dtype))


class _SigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = K.sigmoid_fwd(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return K.sigmoid_bwd(y, dy.contiguous())


def sigmoid(x):
    return _SigmoidFn.apply(x)


class _UpsampleTrilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sf):
        ctx.sf = sf
        return K.upsample_trilinear_fwd(x.contiguous(), sf)

    @staticmethod
    def backward(ctx, dy):
        return K.upsample_trilinear_bwd(dy.contiguous(), ctx.sf), None


def upsample_trilinear(x, scale_factor):
    """1-channel (N,1,D,H,W) fp32 volume, integer scale factor, align_corners=False."""
    if x.shape[1] != 1 or int(scale_factor) != scale_factor:
        raise NotImplementedError("upsample_trilinear: 1-channel volumes and integer scale factors")
    return _UpsampleTrilinearFn.apply(x, int(scale_factor))
