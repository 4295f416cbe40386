"""CPU oracle for the PCRLv2 2-D pre-training path.  TEST INFRASTRUCTURE ONLY.

Restates, in plain functional PyTorch on CPU, what /root/reference/models/pcrlv2_model.py:49-209 and
/root/reference/train_2d.py:111-195 compute.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu legs may import it; nothing under pcrlv2_b200/ does.

Parity status: the 2-D model is built on ``segmentation_models_pytorch`` (absent from
/root/reference and from this image, unpinned: README.md:7).  Its ResNet-18 encoder is torchvision's
``ResNet(BasicBlock, [2,2,2,2])`` minus fc/avgpool, restated below op by op; oracle/make_golden_2d.py
pins this file against (a) torchvision's own resnet18 modules and (b) the UNMODIFIED reference
decoder / model / trainer (``models/pcrlv2_model.py``, ``train_2d.train_pcrlv2_inner``) imported over
oracle/smp_stub (a restatement of the few smp classes the reference touches), and writes
tests/golden/train2d_*.npz.  The reference ships no tests for this path: the pin is the reference's
own code run in the build container, with smp's published structure restated -- stated as such in
DESIGN.md.

State is a flat dict keyed exactly like ``PCRLv2().state_dict()`` of the reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .pcrlv2_oracle import sgd_step, lr_at, clone_state, is_param  # noqa: F401  (same optimiser / schedule)

# models/pcrlv2_model.py:135 (decoder_channels) and :148-156 (in / out channels of the five blocks)
DEC_IN = (512, 256, 128, 64, 32)
DEC_OUT = (256, 128, 64, 32, 16)
# torchvision resnet18: (name, cin, cout, stride of the first block)
LAYERS = (("layer1", 64, 64, 1), ("layer2", 64, 128, 2), ("layer3", 128, 256, 2), ("layer4", 256, 512, 2))


def _bn_spec(spec, prefix, c):
    spec.append((f"{prefix}.weight", (c,), "one"))
    spec.append((f"{prefix}.bias", (c,), "zero"))
    spec.append((f"{prefix}.running_mean", (c,), "zero"))
    spec.append((f"{prefix}.running_var", (c,), "one"))
    spec.append((f"{prefix}.num_batches_tracked", (), "count"))


def state_spec(n_class: int = 3):
    """Ordered (key, shape, kind) list equal to the reference ``PCRLv2().state_dict()`` layout
    (checked against the real thing by make_golden_2d.py)."""
    spec = []
    e = "model.encoder"
    spec.append((f"{e}.conv1.weight", (64, 3, 7, 7), "w"))
    _bn_spec(spec, f"{e}.bn1", 64)
    for name, cin, cout, stride in LAYERS:
        for b in range(2):
            p = f"{e}.{name}.{b}"
            ci = cin if b == 0 else cout
            spec.append((f"{p}.conv1.weight", (cout, ci, 3, 3), "w"))
            _bn_spec(spec, f"{p}.bn1", cout)
            spec.append((f"{p}.conv2.weight", (cout, cout, 3, 3), "w"))
            _bn_spec(spec, f"{p}.bn2", cout)
            if b == 0 and (stride != 1 or cin != cout):
                spec.append((f"{p}.downsample.0.weight", (cout, cin, 1, 1), "w"))
                _bn_spec(spec, f"{p}.downsample.1", cout)
    for i, (ci, co) in enumerate(zip(DEC_IN, DEC_OUT)):
        p = f"model.decoder.blocks.{i}"
        spec.append((f"{p}.conv1.0.weight", (co, ci, 3, 3), "w"))
        _bn_spec(spec, f"{p}.conv1.1", co)
        spec.append((f"{p}.conv2.0.weight", (co, co, 3, 3), "w"))
        _bn_spec(spec, f"{p}.conv2.1", co)
        _bn_spec(spec, f"{p}.bn", co)
        spec.append((f"{p}.deep_supervision_head.0.weight", (co, co, 3, 3), "w"))
        spec.append((f"{p}.deep_supervision_head.0.bias", (co,), f"b:{co * 9}"))
        _bn_spec(spec, f"{p}.deep_supervision_head.1", co)
        spec.append((f"{p}.deep_supervision_head.3.weight", (3, co, 1, 1), "w"))
        spec.append((f"{p}.deep_supervision_head.3.bias", (3,), f"b:{co}"))
        spec.append((f"{p}.predictor_head.0.weight", (2 * co, co), "w"))
        spec.append((f"{p}.predictor_head.0.bias", (2 * co,), f"b:{co}"))
        _bn_spec(spec, f"{p}.predictor_head.1", 2 * co)
        spec.append((f"{p}.predictor_head.3.weight", (co, 2 * co), "w"))
        spec.append((f"{p}.predictor_head.3.bias", (co,), f"b:{2 * co}"))
    spec.append(("model.segmentation_head.0.weight", (n_class, 16, 3, 3), "w"))
    spec.append(("model.segmentation_head.0.bias", (n_class,), "b:144"))
    return spec


def init_state(seed: int = 0, dtype=torch.float32, n_class: int = 3):
    """Deterministic initial state: weights / biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)), norm weight 1 /
    bias 0, running_mean 0 / running_var 1, from a private generator (the GPU box rebuilds the same
    tensors without the reference).  The distributions differ from the reference's initialisers
    (kaiming / xavier, models/pcrlv2_model.py:23-46): parity is checked from a COMMON state."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_spec(n_class):
        if kind == "w" or kind.startswith("b:"):
            if kind == "w":
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
            else:
                fan_in = int(kind[2:])
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        elif kind == "one":
            sd[key] = torch.ones(shape, dtype=dtype)
        elif kind == "zero":
            sd[key] = torch.zeros(shape, dtype=dtype)
        else:
            sd[key] = torch.zeros((), dtype=torch.long)
    return sd


def is_cancelling(key: str) -> bool:
    """Parameters whose exact gradient is zero in train mode because a batch-statistics BatchNorm
    follows directly: the deep-supervision conv bias (Conv2d -> BatchNorm2d, models/pcrlv2_model.py:103-104),
    ``blocks.i.bn.bias`` (-> Linear -> BatchNorm1d, :107-109,125-126; the other use of ``pro`` is detached,
    train_2d.py:115) and ``predictor_head.0.bias`` (Linear -> BatchNorm1d)."""
    return (key.endswith("deep_supervision_head.0.bias") or key.endswith("predictor_head.0.bias")
            or key.endswith(".bn.bias"))


# --------------------------------------------------------------------------- forward
# Every convolution goes through this hook so that oracle/operand_emulation_2d.py can restate the SAME forward
# with the operand rounding of the precision modes (TF32 / bf16 tensor-core operands); the default is F.conv2d.
# kind: "tc" = a convolution the CUDA path runs on the tensor cores, "c3" = one of the 3-output-channel convs.
CONV_IMPL = [lambda x, w, b, stride, pad, kind: F.conv2d(x, w, b, stride, pad)]


def _conv(x, w, b=None, stride=1, pad=0, kind="tc"):
    return CONV_IMPL[0](x, w, b, stride, pad, kind)


def _bn(x, sd, prefix, training):
    """nn.BatchNorm2d / nn.BatchNorm1d (momentum 0.1, eps 1e-5), buffers updated in train mode."""
    if training:
        sd[f"{prefix}.num_batches_tracked"] += 1
    return F.batch_norm(x, sd[f"{prefix}.running_mean"], sd[f"{prefix}.running_var"], sd[f"{prefix}.weight"],
                        sd[f"{prefix}.bias"], training, 0.1, 1e-5)


def _basic_block(x, sd, p, stride, training):
    """torchvision.models.resnet.BasicBlock.forward."""
    out = F.relu(_bn(_conv(x, sd[f"{p}.conv1.weight"], None, stride, 1), sd, f"{p}.bn1", training))
    out = _bn(_conv(out, sd[f"{p}.conv2.weight"], None, 1, 1), sd, f"{p}.bn2", training)
    if f"{p}.downsample.0.weight" in sd:
        x = _bn(_conv(x, sd[f"{p}.downsample.0.weight"], None, stride, 0), sd, f"{p}.downsample.1", training)
    return F.relu(out + x)


def encoder(sd, x, training=True):
    """smp ResNetEncoder.forward for resnet18, depth 5: six features, the first is the input."""
    e = "model.encoder"
    feats = [x]
    x = F.relu(_bn(_conv(x, sd[f"{e}.conv1.weight"], None, 2, 3), sd, f"{e}.bn1", training))
    feats.append(x)
    x = F.max_pool2d(x, 3, 2, 1)
    for name, _cin, _cout, stride in LAYERS:
        x = _basic_block(x, sd, f"{e}.{name}.0", stride, training)
        x = _basic_block(x, sd, f"{e}.{name}.1", 1, training)
        feats.append(x)
    return feats


def decoder_block(x, sd, p, training):
    """DecoderBlock.forward, models/pcrlv2_model.py:113-128 (skip connection commented out :115-117)."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.relu(_bn(_conv(x, sd[f"{p}.conv1.0.weight"], None, 1, 1), sd, f"{p}.conv1.1", training))
    x = F.relu(_bn(_conv(x, sd[f"{p}.conv2.0.weight"], None, 1, 1), sd, f"{p}.conv2.1", training))
    h = f"{p}.deep_supervision_head"
    m = F.relu(_bn(_conv(x, sd[f"{h}.0.weight"], sd[f"{h}.0.bias"], 1, 1), sd, f"{h}.1", training))
    m = _conv(m, sd[f"{h}.3.weight"], sd[f"{h}.3.bias"], 1, 0, "c3")
    pro = _bn(F.adaptive_avg_pool2d(x, (1, 1)).flatten(1), sd, f"{p}.bn", training)
    q = f"{p}.predictor_head"
    pre = F.linear(pro, sd[f"{q}.0.weight"], sd[f"{q}.0.bias"])
    pre = F.relu(_bn(pre, sd, f"{q}.1", training))
    pre = F.linear(pre, sd[f"{q}.3.weight"], sd[f"{q}.3.bias"])
    return x, pro, pre, m


def forward(sd, x, local=False, training=True):
    """PCRLv2.forward, models/pcrlv2_model.py:203-209: ``local`` is NOT forwarded to the decoder (:205),
    so the middle masks are computed and upsampled for local views too; the final mask is skipped."""
    feats = encoder(sd, x, training)
    h = feats[-1]                                               # :177-181: head = deepest feature, skips unused
    outs, masks = [], []
    for i in range(5):
        h, pro, pre, m = decoder_block(h, sd, f"model.decoder.blocks.{i}", training)
        outs.append((pro, pre))
        masks.append(F.interpolate(m, scale_factor=2 ** (4 - i), mode="bilinear"))   # :191-193
    mask = None
    if not local:
        mask = _conv(h, sd["model.segmentation_head.0.weight"], sd["model.segmentation_head.0.bias"], 1, 1, "c3")
    return outs, mask, masks


# --------------------------------------------------------------------------- loss / step
def cos_loss(rng, output1, output2):
    """train_2d.py:111-117."""
    index = rng.randint(0, len(output1) - 1)
    s1, s2 = output1[index], output2[index]
    loss = -(F.cosine_similarity(s1[1], s2[0].detach(), dim=1, eps=1e-8).mean()
             + F.cosine_similarity(s2[1], s1[0].detach(), dim=1, eps=1e-8).mean()) * 0.5
    return loss, index


def step_loss(sd, x1, x2, gt, local_views, epoch, rng):
    """train_2d.py:141-163: three forwards and the four loss terms.  Returns (loss, terms, 13 draws)."""
    bsz = x1.shape[0]
    dec1, mask1, mm1 = forward(sd, x1)
    dec2, _mask2, _ = forward(sd, x2)
    draws = []
    loss2, index2 = cos_loss(rng, dec1, dec2)
    draws.append(index2)
    loc, _, _ = forward(sd, torch.cat(local_views, dim=0), local=True)
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
    loss = loss1 + loss2 + local_loss + loss4
    terms = dict(loss=loss, loss1=loss1, loss2=loss2, loss4=loss4, local_loss=local_loss,
                 mask1=mask1, dec1=dec1, dec2=dec2, mm1=mm1)
    return loss, terms, draws


def train_step(sd, bufs, x1, x2, gt, local_views, epoch, lr, rng, momentum=0.9, weight_decay=1e-4):
    """One iteration of train_2d.train_pcrlv2_inner (:134-173; no step-skip guard in the 2-D trainer).
    Mutates ``sd`` and ``bufs``.  Returns (scalars, draws, grads)."""
    keys = [k for k in sd if is_param(k)]
    for k in keys:
        sd[k].requires_grad_(True)
    loss, terms, draws = step_loss(sd, x1, x2, gt, local_views, epoch, rng)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    for k in keys:
        sd[k].requires_grad_(False)
    grads = dict(zip(keys, grads))
    sgd_step(sd, grads, bufs, lr, momentum, weight_decay)
    scalars = {k: float(terms[k].detach()) for k in ("loss", "loss1", "loss2", "loss4", "local_loss")}
    return scalars, draws, grads


def synthetic_batch(bsz, seed=42, size=(224, 224), local=(96, 96), n_local=6, dtype=torch.float32):
    """Synthetic chest-X-ray-shaped batch (datasets/chestDataset.py:31-48: two global 224^2 crops of one
    PNG, normalised; gt = the first crop before augmentation; six local 96^2 crops)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn((bsz, 3) + tuple(size), generator=g, dtype=dtype)
    x2 = torch.randn((bsz, 3) + tuple(size), generator=g, dtype=dtype)
    gt = torch.rand((bsz, 3) + tuple(size), generator=g, dtype=dtype)
    lv = [torch.randn((bsz, 3) + tuple(local), generator=g, dtype=dtype) for _ in range(n_local)]
    return x1, x2, gt, lv
