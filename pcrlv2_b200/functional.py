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
    def forward(ctx, p, t, weight):
        p, t = p.contiguous(), t.contiguous()
        ctx.save_for_backward(p, t, weight)
        return K.mse_fwd(p, t, weight)

    @staticmethod
    def backward(ctx, g):
        p, t, weight = ctx.saved_tensors
        return K.mse_bwd(p, t, g.contiguous().float(), weight), None, None


def mse_loss(pred, target, weight=None):
    """nn.MSELoss()(pred, target) [* weight: a 1-element device tensor read by the kernels, so a
    captured graph can switch the term on / off and scale it by data]."""
    if target.requires_grad:
        raise NotImplementedError("mse_loss: the target is a constant on this path")
    return _MSEFn.apply(pred, target.to(pred.dtype), weight)


class _ContrastiveFn(torch.autograd.Function):
    """loss2 + local_loss of one step (reference train_3d.py:119,124-133) from the 18 projection /
    prediction tensors, one launch forward+gradient (csrc/losses.cu:contrastive_kernel).

    ``drawn`` (host tuple of the scales that occur in the draws, or None): with it the gradients of
    the prediction tensors of scales NO term drew are returned as ``None`` -- the reference's graph
    does not reach those heads at all and SGD then skips their parameters (SURVEY note N3).  ``None``
    (captured-graph mode): every gradient is returned (zeros where nothing was drawn) and the caller
    supplies the reached-parameter mask itself."""

    @staticmethod
    def forward(ctx, draws, drawn, *tensors):
        S = len(tensors) // 6
        pre1, pro1, pre2, pro2, pre_l, pro_l = (list(tensors[S * i:S * i + S]) for i in range(6))
        out, grads = K.contrastive_fwd_bwd([t.contiguous() for t in pre1], [t.detach().contiguous() for t in pro1],
                                           [t.contiguous() for t in pre2], [t.detach().contiguous() for t in pro2],
                                           [t.contiguous() for t in pre_l], [t.detach().contiguous() for t in pro_l],
                                           draws)
        ctx.drawn, ctx.S = drawn, S
        ctx.save_for_backward(*(grads[0] + grads[1] + grads[2]))
        parts = out.detach().clone()
        ctx.mark_non_differentiable(parts)
        return out.sum(), parts

    @staticmethod
    def backward(ctx, g, _gparts):
        saved, S = ctx.saved_tensors, ctx.S
        dpre1, dpre2, dpre_l = saved[0:S], saved[S:2 * S], saved[2 * S:3 * S]

        def sel(group):
            return [None if (ctx.drawn is not None and s not in ctx.drawn) else group[s] * g for s in range(S)]
        none3 = [None] * S
        return (None, None, *sel(dpre1), *none3, *sel(dpre2), *none3, *sel(dpre_l), *none3)


def contrastive_losses(dec1, dec2, dec_local, draws, drawn=None):
    """dec*: [[pro, pre] x 3 scales] as returned by the model (local: rows = n_local * B, view-major).
    Returns (loss2 + local_loss, parts [2] = (loss2, local_loss))."""
    def pre(dec):
        # a scale that no term drew must not enter the autograd graph at all: the reference never
        # touches its prediction head, so its parameters keep grad None (custom Functions would
        # otherwise be run with materialised zero gradients and mark those parameters as reached)
        return [d[1] if (drawn is None or s in drawn) else d[1].detach() for s, d in enumerate(dec)]
    def pro(dec):      # the reference detaches the projections (train_3d.py:90-91)
        return [d[0].detach() for d in dec]
    args = pre(dec1) + pro(dec1) + pre(dec2) + pro(dec2) + pre(dec_local) + pro(dec_local)
    return _ContrastiveFn.apply(draws, drawn, *args)


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
