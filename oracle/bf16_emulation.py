"""bf16-storage emulation of the oracle.  TEST INFRASTRUCTURE ONLY (see pcrlv2_oracle.py).

The CUDA path keeps activations, raw convolution outputs and activation gradients in bf16 in
HBM (fp32 accumulation, statistics and parameters).  This module restates the reference forward
(models/pcrlv2_model_3d.py:112-133) with round-to-bf16 inserted at exactly those storage points
(forward value AND its gradient), so tests can separate the two contributions to a deviation:

   CUDA  vs  emulation   : kernel correctness (accumulation order, rare ReLU / max-pool flips)
   emulation vs fp32     : the inherent cost of bf16 storage in the reference's own arithmetic
                           (BatchNorm-backward cancellation amplifies it towards the first layers)
"""
import torch
import torch.nn.functional as F

from . import pcrlv2_oracle as orc


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


r = _Round.apply


def _luconv(x, sd, prefix, norm):
    y = r(F.conv3d(x, r(sd[f"{prefix}.conv1.weight"]), None, padding=1))   # bias cancels in the norm
    w, b = sd[f"{prefix}.bn1.weight"], sd[f"{prefix}.bn1.bias"]
    if norm == "bn":
        z = F.batch_norm(y, None, None, w, b, True, 0.1, 1e-5)
    else:
        z = F.instance_norm(y, None, None, w, b, True, 0.1, 1e-5)
    return r(F.relu(z))


def forward(sd, x, local=False, norm="bn"):
    """Train-mode forward with bf16 storage points; BN buffers are not updated."""
    h = x
    for i, name in enumerate(orc.DOWN):
        if i > 0:
            h = F.max_pool3d(h, 2)
        h = _luconv(h, sd, f"{name}.ops.0", norm)
        h = _luconv(h, sd, f"{name}.ops.1", norm)
    feats, masks = [], []
    for name in orc.UP:
        h = r(F.conv_transpose3d(h, r(sd[f"{name}.up_conv.weight"]), sd[f"{name}.up_conv.bias"], stride=2))
        h = _luconv(h, sd, f"{name}.ops.0", norm)
        h = _luconv(h, sd, f"{name}.ops.1", norm)
        pro = F.adaptive_avg_pool3d(h, (1, 1, 1)).view(h.shape[0], -1)
        pro = F.batch_norm(pro, None, None, sd[f"{name}.bn.weight"], sd[f"{name}.bn.bias"], True, 0.1, 1e-5)
        t = F.linear(pro, sd[f"{name}.predictor_head.0.weight"], sd[f"{name}.predictor_head.0.bias"])
        t = F.relu(F.batch_norm(t, None, None, sd[f"{name}.predictor_head.1.weight"],
                                sd[f"{name}.predictor_head.1.bias"], True, 0.1, 1e-5))
        pre = F.linear(t, sd[f"{name}.predictor_head.3.weight"], sd[f"{name}.predictor_head.3.bias"])
        ds = f"{name}.deep_supervision_head"
        y1 = F.conv3d(h, sd[f"{ds}.conv1.weight"], sd[f"{ds}.conv1.bias"], padding=1)
        if norm == "bn":
            z = F.batch_norm(y1, None, None, sd[f"{ds}.bn1.weight"], sd[f"{ds}.bn1.bias"], True, 0.1, 1e-5)
        else:
            z = F.instance_norm(y1, None, None, sd[f"{ds}.bn1.weight"], sd[f"{ds}.bn1.bias"], True, 0.1, 1e-5)
        feats.append([pro, pre])
        masks.append(torch.sigmoid(z))
    mm = []
    if not local:
        mm = [F.interpolate(masks[0], scale_factor=4, mode="trilinear"),
              F.interpolate(masks[1], scale_factor=2, mode="trilinear"), masks[2]]
    out = torch.sigmoid(F.conv3d(h, sd["out_tr.final_conv.weight"], sd["out_tr.final_conv.bias"]))
    return out, feats, mm
