"""B200-native PCRLv2 (2-D): same class surface and state_dict as the reference ``models/pcrlv2_model.py``
(``PCRLv2(n_class=3, low_dim=128)``, ``forward(x, local=False) -> (decoder_outputs, masks, middle_masks)``
:197-209; keys ``model.encoder.*`` / ``model.decoder.blocks.*`` / ``model.segmentation_head.0.*``), different engine.

The reference builds on ``segmentation_models_pytorch.Unet('resnet18')`` and replaces its decoder; smp's
ResNet encoder is torchvision's ``ResNet(BasicBlock, [2, 2, 2, 2])`` without fc / avgpool.  The nn modules below
are PARAMETER CONTAINERS with exactly those attribute names; their ``forward`` is never called.  The arithmetic
runs in libpcrl_b200.so:

  every convolution   im2col (csrc/planar.cu) -> tcgen05 GEMM with the BatchNorm statistics in its epilogue
                      (igemm_kmajor.cu, plain mode) -> norm finalize -> norm+ReLU streaming pass (streaming.cu);
                      backward: two-pass norm/act backward, MN-major split-K GEMM for the weight gradient
                      (igemm_mnmajor.cu), GEMM + gather col2im for the data gradient
  3x3/2 max-pool, residual add+ReLU, nearest x2, bilinear mask upsampling, the 3-channel output convolutions
                      HBM-bound kernels of csrc/planar.cu
  BatchNorm1d / Linear heads     csrc/losses.cu (shared with the 3-D path)

Activations are H-padded NHWC (the 3-D layout with D = 1) in the storage type selected by ``precision``
('fp32' = fp32 storage / TF32 operands, the reference on a GPU; 'bf16' under --amp).  16-channel tensors of the
last decoder block are stored with 32 channels (upper half exactly zero: zero-padded weights, gamma, beta).
A convolution bias in front of a BatchNorm cancels: it is not added, its gradient is exactly zero and it is
folded into running_mean (SURVEY note N1).

First slice of SURVEY 8 f-1: parity first.  Every convolution goes through im2col here (9x the activation
bytes for a 3x3); the implicit-GEMM kernel of the 3-D path needs 2-D spatial tiles before it pays for
224-wide rows.
"""
from __future__ import annotations

import os as _os

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import kernels as K
from .. import kernels2d as K2
from .pcrlv2_model_3d import _PARAM_EPOCH


def _packed2d(conv, cs, dtype, exact=False):
    w = conv.weight
    key = (w._version, _PARAM_EPOCH[0], w.data_ptr(), dtype, cs, exact)
    cache = getattr(conv, "_pcrl_packed2d", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, K2.pack_conv2d_weights(w, cs, dtype, exact))
        conv._pcrl_packed2d = cache
    return cache[1]


def _padv(v, n, fill=0.0):
    """1-D parameter / buffer padded to n entries (detached)."""
    v = v.detach()
    if v.numel() == n:
        return v.contiguous()
    out = torch.full((n,), fill, dtype=v.dtype, device=v.device)
    out[: v.numel()] = v
    return out


class _ConvCfg:
    __slots__ = ("conv", "bn", "k", "s", "p", "act", "training", "dtype", "image", "want_avg", "exact")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


def _conv_stats(x, cfg):
    """im2col -> GEMM (+ per-channel sum / sum of squares) -> BatchNorm scale / shift (running statistics
    updated in train mode).  Returns (y, scale, shift, mean, invstd, dims)."""
    conv, bn, dtype, ex = cfg.conv, cfg.bn, cfg.dtype, bool(cfg.exact)
    cout = conv.weight.shape[0]
    coutp = (cout + 31) // 32 * 32
    col, ho, wo = K2.im2col2d(x, cfg.k, cfg.s, cfg.p, dtype, image=cfg.image, exact=ex)
    cs = x.shape[1] if cfg.image else x.shape[-1]
    n = x.shape[0]
    wmat, _ = _packed2d(conv, cs, dtype, ex)
    bias = _padv(conv.bias, coutp) if conv.bias is not None else None
    gamma, beta = _padv(bn.weight, coutp), _padv(bn.bias, coutp)
    if cfg.training:
        stats = torch.zeros((1, coutp, 2), dtype=torch.float64, device=x.device)
        y = K2.gemm_nt_stats(col, wmat, stats, exact=ex)
        rm, rv = _padv(bn.running_mean, coutp), _padv(bn.running_var, coutp, 1.0)
        scale, shift, mean, invstd = K.norm_finalize(stats, n * ho * wo, gamma, beta, bias, rm, rv,
                                                     bn.num_batches_tracked, 0.1, 1e-5)
        if coutp != cout:           # padded temporaries: copy the real channels back
            bn.running_mean.copy_(rm[:cout])
            bn.running_var.copy_(rv[:cout])
    else:
        y = K2.gemm_nt_any(col, wmat, exact=ex)
        invstd = torch.rsqrt(_padv(bn.running_var, coutp, 1.0) + 1e-5).unsqueeze(0)
        mean = _padv(bn.running_mean, coutp)
        if bias is not None:
            mean = mean - bias
        mean = mean.unsqueeze(0)
        scale = (gamma * invstd).contiguous()
        shift = (beta - mean * scale).contiguous()
        mean, invstd = mean.contiguous(), invstd.contiguous()
    y = y.view(n, 1, ho + 1, wo, coutp)
    return y, scale, shift, mean, invstd, gamma, (n, ho, wo, cout, coutp, cs), col


class _Conv2dBNFn(torch.autograd.Function):
    """Conv2d (any k / stride) -> BatchNorm2d -> ReLU or identity [-> global average].
    Reference: md.Conv2dReLU (models/pcrlv2_model.py:51-64,78-93), torchvision BasicBlock conv-bn pairs,
    the deep-supervision Conv2d(bias)+BatchNorm2d+ReLU (:103-105)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma_, beta_, cfg):
        ctx.set_materialize_grads(False)
        y, scale, shift, mean, invstd, gamma, dims, col = _conv_stats(x, cfg)
        n, ho, wo, cout, coutp, cs = dims
        a, _, avg = K.norm_act_fwd(y, scale, shift, cfg.act, None, want_full=True, want_pool=False,
                                   want_avg=bool(cfg.want_avg), exact=bool(cfg.exact))
        ctx.cfg, ctx.dims = cfg, dims
        ctx.has_bias = bias is not None
        # the im2col matrix is kept for the weight gradient (9x the activation bytes of a 3x3, ~9 GB per b=8 step
        # in fp32: cheaper than a second im2col pass on a 180 GB part); PCRL_2D_SAVE_COL=0 recomputes it instead
        keep = _os.environ.get("PCRL_2D_SAVE_COL", "1") != "0"
        ctx.save_for_backward(x, y, scale, shift, mean, invstd, gamma, col if keep else None)
        if cfg.want_avg:
            return a, (avg[:, :cout] * (1.0 / float(ho * wo))).contiguous()
        return a

    @staticmethod
    def backward(ctx, g_a, g_avg=None):
        cfg = ctx.cfg
        x, y, scale, shift, mean, invstd, gamma, col = ctx.saved_tensors
        n, ho, wo, cout, coutp, cs = ctx.dims
        grads = [None] * 6
        if g_a is None and g_avg is None:
            return tuple(grads)
        gavg = None
        if g_avg is not None:
            gavg = torch.zeros((n, coutp), dtype=torch.float32, device=y.device)
            gavg[:, :cout] = g_avg
        dy, sums = K.norm_act_bwd(y, g_a.contiguous() if g_a is not None else None, None, gavg, scale, shift,
                                  mean, invstd, gamma, cfg.act, None, pool=False, per_sample=False,
                                  batch_stats=cfg.training, exact=bool(cfg.exact))
        sums = sums[0].float()              # one statistics group (BatchNorm); views below, no copies
        grads[3] = sums[:cout, 1]
        grads[4] = sums[:cout, 0]
        if ctx.has_bias:
            grads[2] = torch.zeros(cout, dtype=torch.float32, device=y.device)     # cancels in the BatchNorm
        dy2d = dy.view(n * (ho + 1) * wo, coutp)
        ex = bool(cfg.exact)
        if col is None:
            col, _, _ = K2.im2col2d(x, cfg.k, cfg.s, cfg.p, cfg.dtype, image=cfg.image, exact=ex)
        grads[1] = K2.conv2d_wgrad(dy2d, col, cout, cfg.conv.weight.shape[1], cfg.k, cs, exact=ex)
        del col
        if ctx.needs_input_grad[0]:
            _, wt = _packed2d(cfg.conv, cs, cfg.dtype, ex)
            dcol = K2.gemm_nt_any(dy2d, wt, exact=ex)
            _, h, w, _ = K2.dims2(x)
            grads[0] = K2.col2im2d(dcol, n, h, w, cs, cfg.k, cfg.s, cfg.p, exact=ex)
        return tuple(grads)


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, exact=False):
        ctx.save_for_backward(x)
        ctx.exact = exact
        return K2.maxpool_fwd(x, exact=exact)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return K2.maxpool_bwd(x, g.contiguous(), exact=ctx.exact), None


class _AddReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, exact=False):
        out = K2.add_relu(a, b, 0, exact=exact)
        ctx.save_for_backward(out)
        ctx.exact = exact
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        dg = K2.add_relu(out, g.contiguous(), 1, exact=ctx.exact)
        return dg, dg, None


class _UpNearestFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, exact=False):
        ctx.exact = exact
        return K2.up_nearest_fwd(x, exact=exact)

    @staticmethod
    def backward(ctx, g):
        return K2.up_nearest_bwd(g.contiguous(), exact=ctx.exact), None


class _ConvC3Fn(torch.autograd.Function):
    """Conv2d(C -> 3, k in {1, 3}) + bias -> fp32 NCHW mask."""

    @staticmethod
    def forward(ctx, a, weight, bias, c, exact=False):
        ctx.c, ctx.exact = c, exact
        ctx.save_for_backward(a, weight)
        return K2.conv_c3_fwd(a, weight.detach().contiguous(), bias.detach().contiguous(), c, exact=exact)

    @staticmethod
    def backward(ctx, g):
        a, weight = ctx.saved_tensors
        da, dw, db = K2.conv_c3_bwd(a, weight.detach().contiguous(), g.contiguous(), ctx.c,
                                    need_da=ctx.needs_input_grad[0], exact=ctx.exact)
        return da, dw, db, None, None


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sf):
        ctx.sf = sf
        return K2.bilinear_fwd(x.contiguous(), sf)

    @staticmethod
    def backward(ctx, g):
        return K2.bilinear_bwd(g.contiguous(), ctx.sf), None


# ------------------------------------------------------------------------------ parameter containers
class BasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock (attribute names = state_dict keys)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


class ResNet18Encoder(nn.Module):
    """smp ResNetEncoder('resnet18') = torchvision ResNet(BasicBlock, [2,2,2,2]) minus fc / avgpool."""
    out_channels = (3, 64, 64, 128, 256, 512)

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(BasicBlock(64, 64), BasicBlock(64, 64))
        self.layer2 = nn.Sequential(BasicBlock(64, 128, 2), BasicBlock(128, 128))
        self.layer3 = nn.Sequential(BasicBlock(128, 256, 2), BasicBlock(256, 256))
        self.layer4 = nn.Sequential(BasicBlock(256, 512, 2), BasicBlock(512, 512))
        for m in self.modules():        # torchvision resnet.py initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class Conv2dReLU(nn.Sequential):
    """smp.base.modules.Conv2dReLU with use_batchnorm=True: Conv2d(bias=False), BatchNorm2d, ReLU."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0):
        super().__init__(nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding, bias=False),
                         nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.attention = nn.Identity()


class DecoderBlock(nn.Module):
    """reference :68-128 (skip connection commented out there, :115-117)."""

    def __init__(self, in_channels, skip_channels, out_channels):
        super().__init__()
        self.conv1 = Conv2dReLU(in_channels, out_channels, 3, 1)
        self.attention1 = Attention()
        self.conv2 = Conv2dReLU(out_channels, out_channels, 3, 1)
        self.attention2 = Attention()
        self.bn = nn.BatchNorm1d(out_channels)
        self.deep_supervision_head = nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, padding=1),
                                                   nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
                                                   nn.Conv2d(out_channels, 3, 1))
        self.predictor_head = nn.Sequential(nn.Linear(out_channels, 2 * out_channels),
                                            nn.BatchNorm1d(2 * out_channels), nn.ReLU(inplace=True),
                                            nn.Linear(2 * out_channels, out_channels))


def _initialize_decoder(module):
    """reference :23-37."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class PCRLv2Decoder(nn.Module):
    """reference :131-194."""

    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        in_channels = [enc[0]] + list(decoder_channels[:-1])
        skip_channels = enc[1:] + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList(DecoderBlock(i, s, o) for i, s, o in zip(in_channels, skip_channels, decoder_channels))
        _initialize_decoder(self.blocks)


class _Unet(nn.Module):
    """The three attributes of smp.Unet the reference uses (:200-208)."""

    def __init__(self, n_class):
        super().__init__()
        self.encoder = ResNet18Encoder()
        self.decoder = PCRLv2Decoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, n_class, 3, padding=1), nn.Identity(), nn.Identity())
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)


class PCRLv2(nn.Module):
    """Drop-in for the reference ``PCRLv2`` (:197-209).  ``precision``: 'fp32' (default: fp32 storage / TF32
    operands), 'bf16', or 'fp32x3' (3xTF32 split operands on unrounded fp32 storage: the parity mode)."""

    def __init__(self, n_class=3, low_dim=128, precision="fp32"):
        super().__init__()
        if n_class != 3:
            raise NotImplementedError("the 3-channel output convolutions are specialised for n_class == 3 "
                                      "(the reference default, train_2d.py:65)")
        if precision not in ("fp32", "bf16", "fp32x3"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp32x3'")
        self.model = _Unet(n_class)
        self.precision = precision

    @property
    def _dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    @property
    def _exact(self):
        """precision='fp32x3': fp32 storage without tf32 rounding, every GEMM as three TF32 products of split
        operands (fp32-equivalent products; the parity mode, ~3x the GEMM cost and 3x the im2col bytes)."""
        return self.precision == "fp32x3"

    # ---- pieces
    def _cb(self, x, conv, bn, k, s, p, act="relu", image=False, want_avg=False):
        cfg = _ConvCfg(conv=conv, bn=bn, k=k, s=s, p=p, act=act, training=self.training, dtype=self._dtype,
                       image=image, want_avg=want_avg, exact=self._exact)
        return _Conv2dBNFn.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, cfg)

    def _block(self, x, blk):
        out = self._cb(x, blk.conv1, blk.bn1, 3, blk.stride, 1)
        out = self._cb(out, blk.conv2, blk.bn2, 3, 1, 1, act="none")
        if blk.downsample is not None:
            x = self._cb(x, blk.downsample[0], blk.downsample[1], 1, blk.stride, 0, act="none")
        return _AddReluFn.apply(out, x, self._exact)

    def _encode(self, x):
        enc = self.model.encoder
        h = self._cb(x.float().contiguous(), enc.conv1, enc.bn1, 7, 2, 3, image=True)
        h = _MaxPoolFn.apply(h, self._exact)
        for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4):
            for blk in layer:
                h = self._block(h, blk)
        return h

    def _decode_block(self, h, blk, i, need_masks):
        h = _UpNearestFn.apply(h, self._exact)
        h = self._cb(h, blk.conv1[0], blk.conv1[1], 3, 1, 1)
        h, avg = self._cb(h, blk.conv2[0], blk.conv2[1], 3, 1, 1, want_avg=True)
        x_pro = Fn.batch_norm1d(avg, blk.bn)
        ph = blk.predictor_head
        x_pre = Fn.linear(Fn.batch_norm1d(Fn.linear(x_pro, ph[0]), ph[1], relu=True), ph[3])
        ds = blk.deep_supervision_head
        cout = ds[3].weight.shape[1]
        mask = None
        if need_masks:
            m = self._cb(h, ds[0], ds[1], 3, 1, 1)
            m = _ConvC3Fn.apply(m, ds[3].weight, ds[3].bias, cout, self._exact)
            mask = _BilinearFn.apply(m, 2 ** (4 - i))
        elif self.training:
            # the mask is discarded by the caller, but BatchNorm2d of the head still sees the batch
            # (running statistics, num_batches_tracked): statistics only, no apply pass, no output conv
            with torch.no_grad():
                _conv_stats(h, _ConvCfg(conv=ds[0], bn=ds[1], k=3, s=1, p=1, act="relu", training=True,
                                        dtype=self._dtype, image=False, want_avg=False, exact=self._exact))
        return h, x_pro, x_pre, mask

    def forward(self, x, local=False, need_masks=True):
        """reference :203-209.  ``need_masks=False`` (an extension used by this package's trainer for the
        forwards whose masks the loss never reads, train_2d.py:142,147) skips the output convolutions and
        upsampling of the deep-supervision heads and the segmentation head; their BatchNorm buffers are
        still updated."""
        if x.device.type != "cuda":
            raise RuntimeError("PCRLv2 runs on CUDA (libpcrl_b200.so); there is no CPU path")
        h = self._encode(x)
        outs, middle = [], []
        for i, blk in enumerate(self.model.decoder.blocks):
            h, pro, pre, m = self._decode_block(h, blk, i, need_masks)
            outs.append((pro, pre))
            if m is not None:
                middle.append(m)
        masks = None
        if not local and need_masks:
            seg = self.model.segmentation_head[0]
            masks = _ConvC3Fn.apply(h, seg.weight, seg.bias, seg.weight.shape[1], self._exact)
        return outs, masks, middle


__all__ = ["PCRLv2", "PCRLv2Decoder", "DecoderBlock", "ResNet18Encoder", "BasicBlock", "Conv2dReLU"]
