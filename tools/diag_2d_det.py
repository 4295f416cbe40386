"""Determinism of the 2-D forward: two eager forwards of the same input from the same state, bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcrlv2_oracle_2d as orc
from pcrlv2_b200.models import PCRLv2
from pcrlv2_b200.models import pcrlv2_model as M
m = PCRLv2(precision="fp32"); m.load_state_dict(orc.clone_state(orc.init_state(0))); m = m.cuda().train()
x1 = orc.synthetic_batch(4, seed=50, size=(64, 64), local=(32, 32))[0].cuda()
outs = []
with torch.no_grad():
    for rep in range(3):
        enc = m.model.encoder
        feats = {}
        h = m._cb(x1.float().contiguous(), enc.conv1, enc.bn1, 7, 2, 3, image=True); feats["stem"] = h.clone()
        h = M._MaxPoolFn.apply(h)
        for name in ("layer1", "layer2", "layer3", "layer4"):
            for b in range(2):
                h = m._block(h, getattr(enc, name)[b]); feats[f"{name}.{b}"] = h.clone()
        for i, blk in enumerate(m.model.decoder.blocks):
            h, pro, pre, mask = m._decode_block(h, blk, i, True)
            feats[f"dec{i}.x"] = h.clone(); feats[f"dec{i}.pro"] = pro.clone(); feats[f"dec{i}.mask"] = mask.clone()
        outs.append(feats)
for k in outs[0]:
    d1 = (outs[0][k].float() - outs[1][k].float()).abs().max().item()
    d2 = (outs[0][k].float() - outs[2][k].float()).abs().max().item()
    print(f"{k:12s} max|run0-run1| {d1:.3e}  max|run0-run2| {d2:.3e}  (max |value| {outs[0][k].float().abs().max().item():.3e})")
