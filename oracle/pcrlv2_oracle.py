"""CPU oracle for the PCRLv2 3-D pre-training hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional PyTorch on CPU (fp32, or fp64 when asked), what the
reference computes on the path named by BASELINE.json:north_star.  It exists to CHECK the
CUDA path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.  Nothing under pcrlv2_b200/ imports it.

Parity status: the reference ships no tests or golden vectors, so this
oracle is pinned against the reference itself: oracle/make_golden.py imports
/root/reference/models/pcrlv2_model_3d.py by file path in the build container, checks this
restatement against it (forward outputs, gradients, a 2-step SGD trajectory, BN buffers) and
writes tests/golden/*.npz.  tests/test_oracle_golden.py re-checks the oracle against those
committed fixtures on every run.

Each function cites the reference lines it follows (paths relative to /root/reference).
State is a flat dict keyed exactly like the reference ``PCRLv23d().state_dict()``.
"""
from __future__ import annotations

import math
import random
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- structure
# models/pcrlv2_model_3d.py:37-45 (_make_nConv) + :95-110 (PCRLv23d.__init__)
DOWN = OrderedDict(
    down_tr64=[(None, 32), (32, 64)],   # in_channels filled at build time
    down_tr128=[(64, 64), (64, 128)],
    down_tr256=[(128, 128), (128, 256)],
    down_tr512=[(256, 256), (256, 512)],
)
UP = OrderedDict(
    up_tr256=dict(up=(512, 512), convs=[(512, 256), (256, 256)], ch=256),
    up_tr128=dict(up=(256, 256), convs=[(256, 128), (128, 128)], ch=128),
    up_tr64=dict(up=(128, 128), convs=[(128, 64), (64, 64)], ch=64),
)


def state_spec(in_channels: int = 1, n_class: int = 1, norm: str = "bn", act: str = "relu"):
    """Ordered (key, shape, kind) list equal to the reference state_dict layout.

    kind: 'w' conv/linear weight (fan_in given by shape), 'b:<fan_in>' bias, 'one', 'zero',
    'count' (num_batches_tracked), 'prelu'.
    """
    spec = []

    def luconv(prefix, cin, cout, act_):
        # models/pcrlv2_model_3d.py:6-30
        spec.append((f"{prefix}.conv1.weight", (cout, cin, 3, 3, 3), "w"))
        spec.append((f"{prefix}.conv1.bias", (cout,), f"b:{cin * 27}"))
        spec.append((f"{prefix}.bn1.weight", (cout,), "one"))
        spec.append((f"{prefix}.bn1.bias", (cout,), "zero"))
        if norm == "bn":
            spec.append((f"{prefix}.bn1.running_mean", (cout,), "zero"))
            spec.append((f"{prefix}.bn1.running_var", (cout,), "one"))
            spec.append((f"{prefix}.bn1.num_batches_tracked", (), "count"))
        if act_ == "prelu":
            spec.append((f"{prefix}.activation.weight", (cout,), "prelu"))

    for name, convs in DOWN.items():
        for i, (cin, cout) in enumerate(convs):
            luconv(f"{name}.ops.{i}", in_channels if cin is None else cin, cout, act)
    for name, cfg in UP.items():
        ci, co = cfg["up"]
        # ConvTranspose3d weight is (Cin, Cout, 2,2,2); torch computes fan_in from dim 1
        spec.append((f"{name}.up_conv.weight", (ci, co, 2, 2, 2), "w"))
        spec.append((f"{name}.up_conv.bias", (co,), f"b:{co * 8}"))
        for i, (cin, cout) in enumerate(cfg["convs"]):
            luconv(f"{name}.ops.{i}", cin, cout, act)
        ch = cfg["ch"]
        spec.append((f"{name}.bn.weight", (ch,), "one"))
        spec.append((f"{name}.bn.bias", (ch,), "zero"))
        spec.append((f"{name}.bn.running_mean", (ch,), "zero"))
        spec.append((f"{name}.bn.running_var", (ch,), "one"))
        spec.append((f"{name}.bn.num_batches_tracked", (), "count"))
        spec.append((f"{name}.predictor_head.0.weight", (2 * ch, ch), "w"))
        spec.append((f"{name}.predictor_head.0.bias", (2 * ch,), f"b:{ch}"))
        spec.append((f"{name}.predictor_head.1.weight", (2 * ch,), "one"))
        spec.append((f"{name}.predictor_head.1.bias", (2 * ch,), "zero"))
        spec.append((f"{name}.predictor_head.1.running_mean", (2 * ch,), "zero"))
        spec.append((f"{name}.predictor_head.1.running_var", (2 * ch,), "one"))
        spec.append((f"{name}.predictor_head.1.num_batches_tracked", (), "count"))
        spec.append((f"{name}.predictor_head.3.weight", (ch, 2 * ch), "w"))
        spec.append((f"{name}.predictor_head.3.bias", (ch,), f"b:{2 * ch}"))
        luconv(f"{name}.deep_supervision_head", ch, 1, "sigmoid")
    spec.append(("out_tr.final_conv.weight", (n_class, 64, 1, 1, 1), "w"))
    spec.append(("out_tr.final_conv.bias", (n_class,), "b:64"))
    return spec


def init_state(seed: int = 0, dtype=torch.float32, **kw):
    """Deterministic initial state with torch's default init *distributions*
    (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases;
    norm weight 1 / bias 0; running_mean 0 / running_var 1), drawn from a private generator so
    the GPU box can rebuild the same tensors without the reference present."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_spec(**kw):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        elif kind.startswith("b:"):
            bound = 1.0 / math.sqrt(int(kind[2:]))
            sd[key] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        elif kind == "one":
            sd[key] = torch.ones(shape, dtype=dtype)
        elif kind == "zero":
            sd[key] = torch.zeros(shape, dtype=dtype)
        elif kind == "count":
            sd[key] = torch.zeros((), dtype=torch.long)
        elif kind == "prelu":
            sd[key] = torch.full(shape, 0.25, dtype=dtype)
    return sd


def is_param(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var")
                or key.endswith("num_batches_tracked"))


def is_cancelling(key: str) -> bool:
    """Parameters whose exact gradient is ZERO in train mode because a batch-statistics normalisation
    follows directly (SURVEY note N1): every LUConv ``conv1.bias`` (BatchNorm3d / InstanceNorm3d
    next), ``<up>.bn.bias`` and ``predictor_head.0.bias`` (Linear -> BatchNorm1d next).  The
    reference's gradient for them is rounding noise (1e-15 .. 1e-18 in fp64), so their updates are
    weight decay only and comparisons of their gradients are meaningless."""
    return (key.endswith("conv1.bias") or key.endswith(".bn.bias") or key.endswith("predictor_head.0.bias"))


# --------------------------------------------------------------------------- forward
def _norm3d(x, sd, prefix, norm, training):
    """BatchNorm3d(momentum=0.1, eps=1e-5, affine) / InstanceNorm3d(affine, no running stats)
    models/pcrlv2_model_3d.py:11-16.  Train-mode BN normalises with the biased batch variance
    and stores the unbiased one."""
    w, b = sd[f"{prefix}.weight"], sd[f"{prefix}.bias"]
    if norm == "bn":
        rm, rv = sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"]
        if training:
            sd[f"{prefix}.num_batches_tracked"] += 1
        return F.batch_norm(x, rm, rv, w, b, training, 0.1, 1e-5)
    if norm == "in":
        return F.instance_norm(x, None, None, w, b, True, 0.1, 1e-5)
    raise ValueError(f"normalization type {norm} is not supported")


def _act(x, sd, prefix, act):
    # models/pcrlv2_model_3d.py:20-30
    if act == "relu":
        return F.relu(x)
    if act == "prelu":
        return F.prelu(x, sd[f"{prefix}.activation.weight"])
    if act == "elu":
        return F.elu(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    if act == "leakyrelu":      # not in the reference (extension named by north_star): nn.LeakyReLU(0.01)
        return F.leaky_relu(x, 0.01)
    raise ValueError(f"activation type {act} is not supported")


def luconv(x, sd, prefix, act, norm, training):
    """LUConv.forward, models/pcrlv2_model_3d.py:32-34: act(norm(conv3d_k3p1(x) + b))."""
    y = F.conv3d(x, sd[f"{prefix}.conv1.weight"], sd[f"{prefix}.conv1.bias"], padding=1)
    return _act(_norm3d(y, sd, f"{prefix}.bn1", norm, training), sd, prefix, act)


def _bn1d(x, sd, prefix, training):
    if training:
        sd[f"{prefix}.num_batches_tracked"] += 1
    return F.batch_norm(x, sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"],
                        sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], training, 0.1, 1e-5)


def up_transition(x, sd, name, act, norm, training):
    """UpTransition.forward, models/pcrlv2_model_3d.py:62-72 (no skip concat, :65)."""
    b = x.shape[0]
    x = F.conv_transpose3d(x, sd[f"{name}.up_conv.weight"], sd[f"{name}.up_conv.bias"], stride=2)
    x = luconv(x, sd, f"{name}.ops.0", act, norm, training)
    x = luconv(x, sd, f"{name}.ops.1", act, norm, training)
    pro = F.adaptive_avg_pool3d(x, (1, 1, 1)).view(b, -1)
    pro = _bn1d(pro, sd, f"{name}.bn", training)
    h = F.linear(pro, sd[f"{name}.predictor_head.0.weight"], sd[f"{name}.predictor_head.0.bias"])
    h = F.relu(_bn1d(h, sd, f"{name}.predictor_head.1", training))
    pre = F.linear(h, sd[f"{name}.predictor_head.3.weight"], sd[f"{name}.predictor_head.3.bias"])
    mask = luconv(x, sd, f"{name}.deep_supervision_head", "sigmoid", norm, training)
    return x, pro, pre, mask


def forward(sd, x, local=False, training=True, act="relu", norm="bn"):
    """PCRLv23d.forward, models/pcrlv2_model_3d.py:112-133.  Mutates BN buffers in ``sd``
    exactly like the module does in train mode."""
    h = x
    for i, name in enumerate(DOWN):
        if i > 0:
            h = F.max_pool3d(h, 2)                               # :115-117
        h = luconv(h, sd, f"{name}.ops.0", act, norm, training)
        h = luconv(h, sd, f"{name}.ops.1", act, norm, training)
    feats, masks = [], []
    for name in UP:
        h, pro, pre, m = up_transition(h, sd, name, act, norm, training)
        feats.append([pro, pre])
        masks.append(m)
    middle_masks = []
    if not local:                                                # :124-127
        middle_masks = [F.interpolate(masks[0], scale_factor=4, mode="trilinear"),
                        F.interpolate(masks[1], scale_factor=2, mode="trilinear"),
                        masks[2]]
    out = torch.sigmoid(F.conv3d(h, sd["out_tr.final_conv.weight"], sd["out_tr.final_conv.bias"]))
    return out, feats, middle_masks


# --------------------------------------------------------------------------- loss / step
def cos_loss(rng, output1, output2):
    """train_3d.py:86-92.  ``rng`` is a ``random.Random`` (the reference uses the module-level
    generator; tests seed it explicitly, SURVEY note N4)."""
    index = rng.randint(0, len(output1) - 1)
    s1, s2 = output1[index], output2[index]
    loss = -(F.cosine_similarity(s1[1], s2[0].detach(), dim=1, eps=1e-8).mean()
             + F.cosine_similarity(s2[1], s1[0].detach(), dim=1, eps=1e-8).mean()) * 0.5
    return loss, index


def step_loss(sd, x1, x2, gt, local_views, epoch, rng, act="relu", norm="bn"):
    """train_3d.py:116-138: three forwards and the four loss terms.  Returns
    (loss, dict of terms, list of the 13 drawn scale indices)."""
    bsz = x1.shape[0]
    mask1, dec1, mm1 = forward(sd, x1, False, True, act, norm)
    _mask2, dec2, _ = forward(sd, x2, False, True, act, norm)
    draws = []
    loss2, index2 = cos_loss(rng, dec1, dec2)
    draws.append(index2)
    local_input = torch.cat(local_views, dim=0)
    _, loc, _ = forward(sd, local_input, True, True, act, norm)
    loc = [torch.stack(t) for t in loc]
    local_loss = 0.0
    for i in range(len(local_views)):
        tmp = [t[:, bsz * i: bsz * (i + 1)] for t in loc]
        l1, i1 = cos_loss(rng, dec1, tmp)
        l2, i2 = cos_loss(rng, dec2, tmp)
        draws += [i1, i2]
        local_loss = local_loss + l1 + l2
    local_loss = local_loss / (2 * len(local_views))
    loss1 = F.mse_loss(mask1, gt)
    beta = 0.5 * (1.0 + math.cos(math.pi * epoch / 240))
    loss4 = beta * F.mse_loss(mm1[index2], gt)
    loss = loss1 + loss2 + loss4 + local_loss
    terms = dict(loss=loss, loss1=loss1, loss2=loss2, loss4=loss4, local_loss=local_loss,
                 mask1=mask1, dec1=dec1, dec2=dec2, mm1=mm1)
    return loss, terms, draws


def sgd_step(sd, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD as configured at train_3d.py:48-51; parameters whose grad is None are
    skipped entirely (SURVEY note N3)."""
    with torch.no_grad():
        for k, g in grads.items():
            if g is None:
                continue
            p = sd[k]
            d = g + weight_decay * p
            if k not in bufs:
                bufs[k] = d.clone()
            else:
                bufs[k].mul_(momentum).add_(d)
            p.add_(bufs[k], alpha=-lr)


def lr_at(epoch, lr0, epochs):
    """utils.py:111-114."""
    return lr0 * 0.5 * (1.0 + math.cos(math.pi * epoch / epochs))


def train_step(sd, bufs, x1, x2, gt, local_views, epoch, lr, rng, act="relu", norm="bn",
               momentum=0.9, weight_decay=1e-4):
    """One iteration of train_pcrlv2_inner (train_3d.py:109-151) on CPU.  Mutates ``sd``
    (parameters and BN buffers) and ``bufs`` (momentum).  Returns (terms, draws, grads)."""
    keys = [k for k in sd if is_param(k)]
    for k in keys:
        sd[k].requires_grad_(True)
    loss, terms, draws = step_loss(sd, x1, x2, gt, local_views, epoch, rng, act, norm)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    for k in keys:
        sd[k].requires_grad_(False)
    grads = dict(zip(keys, grads))
    skipped = bool(loss.item() > 1000 and epoch > 10)           # train_3d.py:140-142
    if not skipped:
        sgd_step(sd, grads, bufs, lr, momentum, weight_decay)
    scalars = {k: float(terms[k]) for k in ("loss", "loss1", "loss2", "loss4", "local_loss")}
    return scalars, draws, grads


def synthetic_batch(bsz, seed=42, vol=(64, 64, 32), local=(16, 16, 16), n_local=6,
                    dtype=torch.float32):
    """Synthetic LUNA-shaped batch: x ~ N(0,1) (inputs end in ZNormalization,
    data.py:87), gt ~ U[0,1) (luna_preprocess.py:135-137), 6 local 16^3 views ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn((bsz, 1) + tuple(vol), generator=g, dtype=dtype)
    x2 = torch.randn((bsz, 1) + tuple(vol), generator=g, dtype=dtype)
    gt = torch.rand((bsz, 1) + tuple(vol), generator=g, dtype=dtype)
    lv = [torch.randn((bsz, 1) + tuple(local), generator=g, dtype=dtype) for _ in range(n_local)]
    return x1, x2, gt, lv


def clone_state(sd, dtype=None):
    out = OrderedDict()
    for k, v in sd.items():
        v = v.detach().clone()
        if dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        out[k] = v
    return out


__all__ = ["state_spec", "init_state", "forward", "cos_loss", "step_loss", "sgd_step",
           "train_step", "synthetic_batch", "clone_state", "is_param", "is_cancelling", "lr_at", "random"]
