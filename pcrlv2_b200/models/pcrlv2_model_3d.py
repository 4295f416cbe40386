"""B200-native PCRLv23d: same class surface as the reference ``models/pcrlv2_model_3d.py``
(constructor signature :98, ``forward(x, local=False) -> (out, middle_features, middle_masks)``
:112-133, ``state_dict()`` keys / shapes / dtypes), different engine.

The nn.Conv3d / nn.BatchNorm3d / ... sub-modules below are *parameter containers* only -- they give
the state_dict its reference layout and torch's default initialisation -- their ``forward`` is
never called.  The arithmetic runs in libpcrl_b200.so (hand-written sm_100a kernels) through the
autograd Functions in this file:

  LUConv            conv3d 3x3x3 (tcgen05 implicit GEMM, fused norm statistics)
                    -> norm finalize -> norm+act(+2x2x2 max-pool)(+avg-pool sums) streaming pass
  UpTransition      ConvTranspose3d as a tensor-core GEMM with scatter epilogue, two LUConv, the
                    1-channel deep-supervision conv (and, at the last stage, the 1x1x1 output conv)
  backward          two-pass norm/act/pool backward, conv data gradient (same implicit GEMM with
                    mirrored filters), MN-major split-K weight gradient

Activations between kernels are "H-padded NDHWC" (csrc/common.cuh) in the storage type selected by
``precision`` (fp32 / TF32 operands by default, like the reference on a GPU; bf16 under ``--amp``);
accumulation, norm statistics, parameters and the tensors returned to the caller are fp32.  The convolution bias in
front of a normalisation cancels exactly: it is not added, its gradient is exactly zero, and it
is folded into ``running_mean`` (SURVEY note N1).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from .. import functional as Fn
from .. import kernels as K

# bumped by pcrlv2_b200.train_3d.FlatSGD after every in-place parameter update made through raw
# pointers (torch's version counter does not see those); part of the packed-weight cache key.
_PARAM_EPOCH = [0]


def bump_param_epoch():
    _PARAM_EPOCH[0] += 1


# ---- weight gradients off the critical path ----------------------------------------------------
# In the backward chain  dgrad(L+1) -> norm/act backward(L) -> dgrad(L) -> ...  nothing waits for the
# weight gradient of layer L until the optimizer step.  When the parameters live in a FlatSGD buffer
# the weight-gradient kernels therefore run on a side stream and add their result directly into the
# parameter's slice of the flat gradient buffer: the tensor-core-bound weight-gradient kernel then
# shares the SMs with the HBM-bound norm/act backward passes of the following layers instead of
# queueing behind them.  The main stream re-joins the side stream when backward() finishes (engine
# callback), so ``p.grad`` is complete for whatever runs after ``loss.backward()``.
# PCRL_OVERLAP_WGRAD=0 disables it.
# Limitation (documented): on this path the 3x3x3 weight gradient is accumulated in
# place and autograd sees ``None`` for it, so ``torch.autograd.grad(loss, conv_weight)`` and tensor
# hooks on those weights do not observe it; use ``loss.backward()`` (what the trainer does) or set
# PCRL_OVERLAP_WGRAD=0.
import os as _os
import types as _types

_SIDE = {}            # device index -> side stream
_SIDE_DIRTY = set()   # device indices whose side stream has weight-gradient work queued since the last join
_JOIN_PENDING = [None]    # id of the autograd graph task that already queued the join callback
_EXPLICIT_JOIN = [False]  # captured-graph step: the caller joins the side stream itself (FlatSGD.step_static)


def _side_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE.get(idx)
    if st is None:
        st = _SIDE[idx] = torch.cuda.Stream(device=idx)
    return st


def join_side_streams():
    """Make the current stream wait for the weight-gradient work queued on the side streams.  Streams with
    nothing queued since the last join are left alone: waiting on an idle stream that is not part of an ongoing
    CUDA-graph capture (the 2-D path never forks onto it) would invalidate the capture."""
    for idx, st in _SIDE.items():
        if idx in _SIDE_DIRTY:
            torch.cuda.current_stream(st.device).wait_stream(st)
    _SIDE_DIRTY.clear()
    _JOIN_PENDING[0] = None


def _overlap_target(weight):
    """(optimizer, index) when the weight gradient may be accumulated in place from the side stream."""
    if _os.environ.get("PCRL_OVERLAP_WGRAD", "1") == "0":
        return None
    flat = getattr(weight, "_pcrl_flat", None)
    if flat is None or weight.grad is None or not weight.grad.is_contiguous():
        return None
    return flat


def _wgrad_overlapped(dy, x, weight, flat, exact=False):
    side = _side_stream(dy.device)
    _SIDE_DIRTY.add(dy.device.index if dy.device.index is not None else torch.cuda.current_device())
    side.wait_stream(torch.cuda.current_stream())          # dy (and zero_grad) are ordered before
    with torch.cuda.stream(side):
        g = K.unpack_conv3_wgrad(K.conv3d_k3_wgrad(dy, x, exact=exact))
        weight.grad.add_(g)
    # dy / x were allocated on the main stream: the allocator may not hand their memory out again
    # before the side stream is done with them
    dy.record_stream(side)
    x.record_stream(side)
    opt, idx = flat
    opt._touched[idx] = True
    # one join per backward(): keyed on the running graph task, so a backward that raised before its
    # callback ran cannot leave the flag stuck for the next one
    task = torch._C._current_graph_task_id()
    if not _EXPLICIT_JOIN[0] and _JOIN_PENDING[0] != task:
        _JOIN_PENDING[0] = task
        torch.autograd.Variable._execution_engine.queue_callback(join_side_streams)


def _packed(module, kind, dtype=torch.bfloat16, exact=False):
    """Tensor-core operand layouts of a conv weight in the activation storage type, cached until
    the parameter changes."""
    w = module.weight
    key = (w._version, _PARAM_EPOCH[0], w.data_ptr(), dtype, exact)
    cache = getattr(module, "_pcrl_packed", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            if kind == "conv3":
                pk = K.pack_conv3_weights(w.detach().contiguous(), dtype=dtype, exact=exact)
            elif kind == "convT":
                pk = K.pack_convT_weights(w.detach().contiguous(), dtype=dtype, exact=exact)
            elif kind == "head":    # (ds conv, final conv or None) -> wext [32,C], wextT [C,32]
                raise ValueError("use _packed_head")
            else:
                raise ValueError(kind)
        cache = (key, pk)
        module._pcrl_packed = cache
    return cache[1]


def _packed_head(ds, fin, dtype=torch.bfloat16, exact=False):
    """bf16 GEMM operands of the 1-channel head convolutions (deep-supervision conv [+ 1x1x1 output
    conv]), cached until either weight changes."""
    key = (ds.weight._version, ds.weight.data_ptr(), _PARAM_EPOCH[0], dtype, exact,
           None if fin is None else (fin.weight._version, fin.weight.data_ptr()))
    cache = getattr(ds, "_pcrl_head", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            pk = K.head_pack_weights(ds.weight.detach(), None if fin is None else fin.weight.detach(), dtype=dtype,
                                     exact=exact)
        cache = (key, pk)
        ds._pcrl_head = cache
    return cache[1]


class _Cfg:
    """Static (non-tensor) configuration of one fused LUConv call."""
    __slots__ = ("stem", "pool", "tail", "final", "act", "norm", "training", "conv", "bn", "ds", "fin", "up", "dtype",
                 "exact")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class _LUConvFn(torch.autograd.Function):
    """[ConvTranspose3d(k2,s2) ->] conv3x3x3 -> norm -> act [-> maxpool]
    [-> avg-pool sums, 1-channel head convs].

    With ``cfg.up`` set the ConvTranspose that feeds the convolution is part of the same node, so
    the backward pass can have the data-gradient kernel write straight into the coarse-major
    layout the ConvTranspose gradient GEMMs read (no re-layout pass over the largest tensors)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, prelu, ds_w, ds_b, fin_w, fin_b, up_w, up_b, cfg):
        ctx.set_materialize_grads(False)
        ex = bool(cfg.exact)
        per_sample = cfg.norm == "in"
        cout = weight.shape[0]
        use_batch_stats = cfg.training or per_sample
        x_coarse = None
        if cfg.up is not None:
            wtf, _ = _packed(cfg.up, "convT", cfg.dtype, ex)
            x_coarse = x
            x = K.convT_fprop(x, wtf, up_b.detach().contiguous(), exact=ex)
        if cfg.stem:
            n, _, d, h, w = x.shape
        else:
            n, d, h, w, _ = K.dims_of(x)
        groups = n if per_sample else 1
        stats = (torch.zeros((groups, cout, 2), dtype=torch.float64, device=x.device)
                 if use_batch_stats else None)
        x27 = None
        if cfg.stem and use_batch_stats and not per_sample and cfg.dtype == torch.bfloat16:
            # bf16 mode: Conv3d(1 -> 32) as im2col (27 taps -> 32 columns) + a K = 32 tensor-core GEMM with
            # the statistics epilogue: the copy rate instead of 15 % of it; X27 is kept for the weight
            # gradient.  (fp32 mode keeps the exact-fp32 SIMT stem: rounding the network INPUT to tf32
            # moved `out` from 1.04e-3 to 1.11e-3 of the fp32 reference for 1.2 % of the step.)
            x27 = K.im2col27(x.contiguous(), cfg.dtype)
            y = K.stem_conv_fprop_gemm(x27, K.stem_pack_weights(weight, cfg.dtype), (n, d, h, w), stats)
        elif cfg.stem:
            y = K.stem_conv_fprop(x.contiguous(), weight.detach().contiguous(), stats, per_sample, dtype=cfg.dtype,
                                  exact=ex)
        else:
            wf, _ = _packed(cfg.conv, "conv3", cfg.dtype, ex)
            y = K.conv3d_k3_fprop(x, wf, stats, per_sample, exact=ex)
        if use_batch_stats:
            count = d * h * w * (1 if per_sample else n)
            bn = cfg.bn
            track = (not per_sample) and cfg.training
            scale, shift, mean, invstd = K.norm_finalize(
                stats, count, gamma.detach(), beta.detach(), bias.detach(),
                bn.running_mean if track else None, bn.running_var if track else None,
                bn.num_batches_tracked if track else None, 0.1, 1e-5)
        else:
            bn = cfg.bn
            invstd = torch.rsqrt(bn.running_var + 1e-5).unsqueeze(0)
            mean = (bn.running_mean - bias.detach()).unsqueeze(0)
            scale = (gamma.detach() * invstd).contiguous()
            shift = (beta.detach() - mean * scale).contiguous()
            mean, invstd = mean.contiguous(), invstd.contiguous()
        slope = prelu.detach() if prelu is not None else None
        a, pooled, avg = K.norm_act_fwd(y, scale, shift, cfg.act, slope, want_full=not cfg.pool,
                                        want_pool=cfg.pool, want_avg=cfg.tail, per_sample=per_sample, exact=ex)
        outs = [pooled if cfg.pool else a]
        if cfg.tail:
            wext, _ = _packed_head(cfg.ds, cfg.fin, cfg.dtype, ex)
            st1 = torch.zeros((groups, 1, 2), dtype=torch.float64, device=x.device)
            y1, y0 = K.head_fwd(a, wext, ds_b.detach(), fin_b.detach() if cfg.final else None,
                                st1, per_sample, exact=ex)
            # the kernel accumulates sums; hand autograd the MEAN so that the incoming gradient is
            # dL/d(mean), which is what norm_act_bwd expects for its gavg argument
            outs += [avg * (1.0 / float(d * h * w)), y1, st1]
            if cfg.final:
                outs.append(y0)
            ctx.mark_non_differentiable(st1)
        ctx.cfg = cfg
        ctx.dims = (n, d, h, w, cout)
        ctx.save_for_backward(x if x27 is None else x27, y, scale, shift, mean, invstd, gamma, prelu,
                              a if cfg.tail else None, ds_w, fin_w, x_coarse)
        ctx.stem_x27 = x27 is not None
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_out, g_avg=None, g_y1=None, _g_st1=None, g_y0=None):
        cfg = ctx.cfg
        ex = bool(cfg.exact)
        x, y, scale, shift, mean, invstd, gamma, prelu, a, ds_w, fin_w, x_coarse = ctx.saved_tensors
        n, d, h, w, cout = ctx.dims
        per_sample = cfg.norm == "in"
        grads = [None] * 13
        g2 = None
        if cfg.tail and (g_y1 is not None or g_y0 is not None):
            dy1 = g_y1.contiguous() if g_y1 is not None else torch.zeros(
                (n, 1, d, h, w), dtype=torch.float32, device=y.device)
            dy0 = g_y0.contiguous() if (cfg.final and g_y0 is not None) else None
            _, wext_t = _packed_head(cfg.ds, cfg.fin, cfg.dtype, ex)
            g2, dwext = K.head_bwd(a, dy1, dy0, wext_t, exact=ex)
            if g_y1 is not None:
                grads[6] = dwext[:, :27].reshape(1, cout, 3, 3, 3)
                grads[7] = dy1.sum().reshape(1)
            if dy0 is not None:
                grads[8] = dwext[:, 27].reshape(1, cout, 1, 1, 1)
                grads[9] = dy0.sum().reshape(1)
        if g_out is None and g2 is None and g_avg is None:
            return tuple(grads)
        g1 = g_out.contiguous() if g_out is not None else None
        gavg = g_avg.contiguous() if g_avg is not None else None
        dy, sums = K.norm_act_bwd(y, g1, g2, gavg, scale, shift, mean, invstd, gamma.detach(), cfg.act,
                                  prelu.detach() if prelu is not None else None, pool=cfg.pool,
                                  per_sample=per_sample, batch_stats=cfg.training or per_sample, exact=ex)
        sums = sums.sum(0).float()
        grads[3] = sums[:, 1].contiguous()          # d gamma
        grads[4] = sums[:, 0].contiguous()          # d beta
        if prelu is not None:
            grads[5] = sums[:, 2].contiguous()
        grads[2] = torch.zeros(cout, dtype=torch.float32, device=y.device)  # conv bias: exactly 0
        if cfg.stem:
            grads[1] = (K.stem_conv_wgrad_gemm(dy, None, x27=x) if ctx.stem_x27
                        else K.stem_conv_wgrad_gemm(dy, x, exact=ex))
        else:
            _, wd = _packed(cfg.conv, "conv3", cfg.dtype, ex)
            flat = _overlap_target(cfg.conv.weight)
            if flat is not None:
                _wgrad_overlapped(dy, x, cfg.conv.weight, flat, ex)  # grads[1] stays None: added in place
            elif cout % 64:
                # 32 output channels on the tensor-core path only occur for the multi-channel stem: the weight-
                # gradient kernel wants a 64-row operand, so dY is zero-padded for this one call
                gpk = K.conv3d_k3_wgrad(torch.nn.functional.pad(dy, (0, 64 - cout % 64)), x, exact=ex)
                grads[1] = K.unpack_conv3_wgrad(gpk)[:cout].contiguous()
            else:
                grads[1] = K.unpack_conv3_wgrad(K.conv3d_k3_wgrad(dy, x, exact=ex))
            if cfg.up is not None:
                # data gradient lands coarse-major; its column sums are the ConvTranspose bias gradient
                scratch, colsum = K.conv3d_k3_dgrad_unshuffled(dy, wd, exact=ex)
                _, wtd = _packed(cfg.up, "convT", cfg.dtype, ex)
                dxc, dwt = K.convT_bwd_from_scratch(scratch, x_coarse, wtd, need_dx=ctx.needs_input_grad[0], exact=ex)
                cin_t, cout_t = cfg.up.weight.shape[0], cfg.up.weight.shape[1]
                grads[0] = dxc
                grads[10] = K.unpack_convT_wgrad(dwt, cin_t, cout_t)
                grads[11] = colsum[:, 0].float().contiguous()
            elif ctx.needs_input_grad[0]:
                grads[0] = K.conv3d_k3_dgrad(dy, wd, exact=ex)
        return tuple(grads)


class _Chan1NormSigmoidFn(torch.autograd.Function):
    """mask = sigmoid(BatchNorm3d(1) / InstanceNorm3d(1) (y1)) of a deep-supervision head
    (reference :12,27,71).  ``stats`` are the (sum, sum of squares) the head gather produced."""

    @staticmethod
    def forward(ctx, y1, gamma, beta, stats, bn, norm, training):
        per_sample = norm == "in"
        n = y1.shape[0]
        vol = y1.numel() // n
        if training or per_sample:
            track = training and not per_sample
            scale, shift, mean, invstd = K.norm_finalize(
                stats, vol if per_sample else y1.numel(), gamma.detach(), beta.detach(), None,
                bn.running_mean if track else None, bn.running_var if track else None,
                bn.num_batches_tracked if track else None, 0.1, 1e-5)
        else:
            invstd = torch.rsqrt(bn.running_var + 1e-5).reshape(1, 1)
            mean = bn.running_mean.reshape(1, 1).clone()
            scale = (gamma.detach().reshape(1, 1) * invstd).contiguous()
            shift = (beta.detach().reshape(1, 1) - mean * scale).contiguous()
        mask = K.chan1_sigmoid_fwd(y1, scale, shift, per_sample)
        ctx.per_sample = per_sample
        ctx.batch_stats = training or per_sample
        ctx.save_for_backward(y1, mask, mean, invstd, gamma)
        return mask

    @staticmethod
    def backward(ctx, dmask):
        y1, mask, mean, invstd, gamma = ctx.saved_tensors
        dy, sums = K.chan1_sigmoid_bwd(y1, mask, dmask.contiguous(), mean, invstd, gamma.detach(),
                                       ctx.per_sample, batch_stats=ctx.batch_stats)
        sums = sums.sum(0).float()
        return dy, sums[1].reshape(1), sums[0].reshape(1), None, None, None, None


class _ConvTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        wf, _ = _packed(mod, "convT")
        ctx.mod = mod
        ctx.save_for_backward(x)
        ctx.shape = tuple(weight.shape)
        return K.convT_fprop(x, wf, bias.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        _, wd = _packed(ctx.mod, "convT")
        cin, cout = ctx.shape[0], ctx.shape[1]
        dx, dw, db = K.convT_bwd(g.contiguous(), x, wd, need_dx=ctx.needs_input_grad[0])
        return dx, K.unpack_convT_wgrad(dw, cin, cout), db, None


def _make_norm(norm, ch):
    if norm == "bn":
        return nn.BatchNorm3d(num_features=ch, momentum=0.1, affine=True)
    if norm == "gn":
        return nn.GroupNorm(num_groups=8, num_channels=ch, eps=1e-05, affine=True)
    if norm == "in":
        return nn.InstanceNorm3d(num_features=ch, momentum=0.1, affine=True)
    raise ValueError("normalization type {} is not supported".format(norm))


class LUConv(nn.Module):
    """Conv3d(k3,p1) -> norm -> activation (reference :6-34)."""

    def __init__(self, in_chan, out_chan, act, norm):
        super().__init__()
        self.conv1 = nn.Conv3d(in_chan, out_chan, kernel_size=3, padding=1)
        self.bn1 = _make_norm(norm, out_chan)
        if norm == "gn":
            raise NotImplementedError(
                "norm='gn' cannot be built by the reference either (GroupNorm(8) on the 1-channel "
                "deep-supervision head raises); not supported")
        if act == "relu":
            self.activation = nn.ReLU(inplace=True)
        elif act == "prelu":
            self.activation = nn.PReLU(out_chan)
        elif act == "elu":
            self.activation = nn.ELU(inplace=True)
        elif act == "sigmoid":
            self.activation = nn.Sigmoid()
        elif act == "leakyrelu":
            # extension: the reference's LUConv has no LeakyReLU branch (:20-27); north_star names the
            # Conv3d+InstanceNorm+LeakyReLU block (use with norm='in')
            self.activation = nn.LeakyReLU(0.01, inplace=True)
        else:
            raise ValueError("activation type {} is not supported".format(act))
        self.act, self.norm = act, norm
        self.in_chan, self.out_chan = in_chan, out_chan

    def run(self, x, pool=False, tail=None, final=None, up=None, dtype=torch.bfloat16, exact=False):
        """x: fp32 (N,1,D,H,W) for the stem, otherwise an H-padded activation in ``dtype``.
        ``up``: the ConvTranspose3d module to apply to x first (UpTransition).
        ``exact``: 3xTF32 split operands on unrounded fp32 storage (precision='fp32x3')."""
        conv, weight = self.conv1, self.conv1.weight
        if self.in_chan != 1 and self.in_chan < 32 and x.dim() == 5 and x.shape[1] == self.in_chan:
            # multi-channel network input (in_channels > 1, reference :98,104): NCDHW fp32 -> H-padded NDHWC with
            # the channels zero-padded to 32, the weight zero-padded likewise (autograd slices its gradient back);
            # from here on it is an ordinary 32-channel layer of the tensor-core path
            xp = torch.nn.functional.pad(K.pad_ndhwc(x, torch.float32), (0, 32 - self.in_chan))
            if dtype == torch.bfloat16:
                xp = xp.to(dtype)
            elif not exact:      # fp32 storage holds tf32-rounded operands (cvt.rna on the magnitude bits)
                xp = ((xp.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
            x = xp.contiguous()
            weight = torch.nn.functional.pad(self.conv1.weight, (0, 0, 0, 0, 0, 0, 0, 32 - self.in_chan))
            conv = _types.SimpleNamespace(weight=weight)
        cfg = _Cfg(stem=self.in_chan == 1, pool=pool, tail=tail is not None, final=final is not None,
                   act=self.act, norm=self.norm, training=self.training, conv=conv, bn=self.bn1,
                   ds=tail.conv1 if tail is not None else None, fin=final, up=up, dtype=dtype, exact=exact)
        prelu = self.activation.weight if self.act == "prelu" else None
        return _LUConvFn.apply(
            x, weight, self.conv1.bias, self.bn1.weight, self.bn1.bias, prelu,
            tail.conv1.weight if tail is not None else None,
            tail.conv1.bias if tail is not None else None,
            final.weight if final is not None else None,
            final.bias if final is not None else None,
            up.weight if up is not None else None,
            up.bias if up is not None else None, cfg)

    def forward(self, x):
        """Stand-alone use with the reference's NCDHW fp32 convention."""
        inp = x.float() if self.in_chan == 1 else K.pad_ndhwc(x)  # bf16 storage
        if self.in_chan != 1 and self.in_chan % 32:
            raise NotImplementedError("LUConv needs in_chan == 1 or a multiple of 32")
        return K.unpad_ndhwc(self.run(inp)[0])


def _make_nConv(in_channel, depth, act, norm, double_chnnel=False):
    if double_chnnel:
        layer1 = LUConv(in_channel, 32 * (2 ** (depth + 1)), act, norm)
        layer2 = LUConv(32 * (2 ** (depth + 1)), 32 * (2 ** (depth + 1)), act, norm)
    else:
        layer1 = LUConv(in_channel, 32 * (2 ** depth), act, norm)
        layer2 = LUConv(32 * (2 ** depth), 32 * (2 ** depth) * 2, act, norm)
    return nn.Sequential(layer1, layer2)


class DownTransition(nn.Module):
    def __init__(self, in_channel, depth, act, norm):
        super().__init__()
        self.ops = _make_nConv(in_channel, depth, act, norm)

    def run(self, x, pool, dtype=torch.bfloat16, exact=False):
        return self.ops[1].run(self.ops[0].run(x, dtype=dtype, exact=exact)[0], pool=pool, dtype=dtype, exact=exact)[0]


class UpTransition(nn.Module):
    """reference :48-72 (the skip concatenation is commented out there, :65)."""

    def __init__(self, inChans, outChans, depth, act, norm):
        super().__init__()
        self.depth = depth
        self.up_conv = nn.ConvTranspose3d(inChans, outChans, kernel_size=2, stride=2)
        self.ops = _make_nConv(outChans, depth, act, norm, double_chnnel=True)
        channels = 32 * (2 ** depth) * 2
        self.bn = nn.BatchNorm1d(channels)
        self.predictor_head = nn.Sequential(nn.Linear(channels, 2 * channels),
                                            nn.BatchNorm1d(2 * channels),
                                            nn.ReLU(inplace=True),
                                            nn.Linear(2 * channels, channels))
        self.deep_supervision_head = LUConv(channels, 1, "sigmoid", norm)
        self.norm = norm

    def run(self, x, final=None, dtype=torch.bfloat16, exact=False):
        h = self.ops[0].run(x, up=self.up_conv, dtype=dtype, exact=exact)[0]
        outs = self.ops[1].run(h, tail=self.deep_supervision_head, final=final, dtype=dtype, exact=exact)
        a, avg, y1, st1 = outs[0], outs[1], outs[2], outs[3]
        y0 = outs[4] if final is not None else None
        # projection (BatchNorm1d) and prediction (Linear-BN-ReLU-Linear) heads, reference :54-58,67-69:
        # the nn modules hold the parameters / running statistics, csrc/losses.cu does the arithmetic
        x_pro = Fn.batch_norm1d(avg, self.bn)
        ph = self.predictor_head
        x_pre = Fn.linear(Fn.batch_norm1d(Fn.linear(x_pro, ph[0]), ph[1], relu=True), ph[3])
        bn = self.deep_supervision_head.bn1
        mask = _Chan1NormSigmoidFn.apply(y1, bn.weight, bn.bias, st1, bn, self.norm, self.training)
        return a, x_pro, x_pre, mask, y0


class _FinalConvNFn(torch.autograd.Function):
    """Conv3d(64 -> n_class > 1, kernel 1) as a tensor-core GEMM over the rows of the H-padded activation
    (n_class == 1, the reference default, rides as a column of the deep-supervision head GEMM instead)."""

    @staticmethod
    def forward(ctx, a, weight, bias, exact):
        n, d, h, w, c = K.dims_of(a)
        k = weight.shape[0]
        if k > 32:
            raise NotImplementedError("n_class <= 32")
        wp = torch.zeros((32, c), dtype=torch.float32, device=a.device)
        wp[:k] = weight.detach().reshape(k, c)
        bp = torch.zeros((32,), dtype=torch.float32, device=a.device)
        bp[:k] = bias.detach()
        a2d = a.view(-1, c)
        if exact:
            y = torch.empty((a2d.shape[0], 32), dtype=torch.float32, device=a.device)
            _lib.call("pcrl_gemm_nt", K.split3(a2d, 0), K.split3(wp, 1), y, bp, a2d.shape[0], 3 * c, 32, 32, 1, K.F32X)
        else:
            y = K.gemm_nt(a2d, K2_round(wp, a.dtype), bp, out_fp32=True)
        ctx.save_for_backward(a, wp)
        ctx.k, ctx.exact = k, exact
        return y.view(n, d, h + 1, w, 32)[:, :, 1:, :, :k].permute(0, 4, 1, 2, 3).contiguous()

    @staticmethod
    def backward(ctx, g):
        a, wp = ctx.saved_tensors
        n, d, h, w, c = K.dims_of(a)
        k = ctx.k
        g2 = torch.zeros((n, d, h + 1, w, 32), dtype=a.dtype, device=a.device)
        g2[:, :, 1:, :, :k] = g.permute(0, 2, 3, 4, 1)
        g2d, a2d = g2.view(-1, 32), a.view(-1, c)
        wt = wp.t().contiguous()                       # [64][32]
        if ctx.exact:
            da = torch.empty_like(a)
            _lib.call("pcrl_gemm_nt", K.split3(g2d, 0), K.split3(wt, 1), da, None, g2d.shape[0], 96, c, c, 1, K.F32X)
            dwt = torch.zeros((c, 32), dtype=torch.float32, device=a.device)
            _lib.call("pcrl_gemm_tn", K.split3(a2d, 1, stack=True), K.split3(g2d, 0, stack=True), dwt,
                      3 * a2d.shape[0], c, 32, K.F32X)
        else:
            da = K.gemm_nt(g2d, K2_round(wt, a.dtype), out_fp32=False).view_as(a)
            dwt = K.gemm_tn(a2d, g2d)
        dw = dwt.t()[:k].reshape(k, c, 1, 1, 1).contiguous()
        return da, dw, g.sum((0, 2, 3, 4)), None


def K2_round(w32, dtype):
    """fp32 weights -> GEMM operand in the storage type (fp32: cvt.rna.tf32 of the magnitude bits)."""
    if dtype == torch.float32:
        return ((w32.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return w32.to(dtype).contiguous()


class OutputTransition(nn.Module):
    def __init__(self, inChans, n_labels):
        super().__init__()
        self.final_conv = nn.Conv3d(inChans, n_labels, kernel_size=1)
        self.sigmoid = nn.Sigmoid()


class PCRLv23d(nn.Module):
    def __init__(self, n_class=1, act="relu", norm="bn", in_channels=1, low_dim=128, student=False,
                 precision="fp32"):
        """Same arguments as the reference (:98) plus ``precision``: the storage type of the
        activations between kernels.  ``"fp32"`` (default, = the reference's default): fp32 storage,
        TF32 tensor-core operands -- what the reference itself runs on an Ampere-or-newer GPU (torch's
        default ``cudnn.allow_tf32``).  ``"bf16"``: bf16 storage, bf16 tensor-core operands (what
        ``--amp`` selects, reference train_3d.py:52-53).  ``"fp32x3"``: fp32 storage without the tf32
        rounding and every tensor-core product evaluated as three TF32 products of split operands
        (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo): fp32-equivalent arithmetic -- what the reference computes
        with ``allow_tf32=False`` -- at about a third of the speed; the mode in which parity with the
        fp32 reference is asserted to 1e-3 (tests/test_step_gpu.py).  fp32 accumulation, statistics
        and parameters in all three."""
        super().__init__()
        if precision not in ("bf16", "fp32", "fp32x3"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp32x3'")
        self.precision = precision
        if not (1 <= in_channels < 32) or not (1 <= n_class <= 32):
            raise NotImplementedError("in_channels must be in 1..31 and n_class in 1..32")
        self.in_channels, self.n_class = in_channels, n_class
        self.maxpool = nn.MaxPool3d(2)
        self.down_tr64 = DownTransition(in_channels, 0, act, norm)
        self.down_tr128 = DownTransition(64, 1, act, norm)
        self.down_tr256 = DownTransition(128, 2, act, norm)
        self.down_tr512 = DownTransition(256, 3, act, norm)
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.up_tr256 = UpTransition(512, 512, 2, act, norm)
        self.up_tr128 = UpTransition(256, 256, 1, act, norm)
        self.up_tr64 = UpTransition(128, 128, 0, act, norm)
        self.out_tr = OutputTransition(64, n_class)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x, local=False):
        if not x.is_cuda:
            raise RuntimeError("pcrlv2_b200.PCRLv23d runs on CUDA only (there is no CPU fallback)")
        if x.dim() != 5 or x.shape[1] != self.in_channels or any(s % 8 for s in x.shape[2:]):
            raise ValueError("expected (B,%d,D,H,W) with D,H,W multiples of 8, got %s" % (self.in_channels, tuple(x.shape)))
        x = x.float().contiguous()
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        ex = self.precision == "fp32x3"
        h = self.down_tr64.run(x, pool=True, dtype=dt, exact=ex)
        h = self.down_tr128.run(h, pool=True, dtype=dt, exact=ex)
        h = self.down_tr256.run(h, pool=True, dtype=dt, exact=ex)
        h = self.down_tr512.run(h, pool=False, dtype=dt, exact=ex)
        h, pro_256, pre_256, m256, _ = self.up_tr256.run(h, dtype=dt, exact=ex)
        h, pro_128, pre_128, m128, _ = self.up_tr128.run(h, dtype=dt, exact=ex)
        if self.n_class == 1:      # the 1x1x1 output conv rides in the deep-supervision head GEMM
            h, pro_64, pre_64, m64, y0 = self.up_tr64.run(h, final=self.out_tr.final_conv, dtype=dt, exact=ex)
        else:
            h, pro_64, pre_64, m64, _ = self.up_tr64.run(h, dtype=dt, exact=ex)
            fc = self.out_tr.final_conv
            y0 = _FinalConvNFn.apply(h, fc.weight, fc.bias, ex)
        middle_masks = []
        if not local:
            middle_masks.append(Fn.upsample_trilinear(m256, 4))
            middle_masks.append(Fn.upsample_trilinear(m128, 2))
            middle_masks.append(m64)
        middle_features = [[pro_256, pre_256], [pro_128, pre_128], [pro_64, pre_64]]
        out = Fn.sigmoid(y0)
        return out, middle_features, middle_masks
