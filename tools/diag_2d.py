"""Stage-by-stage comparison of the 2-D model (CUDA) with the CPU oracle: where does the error enter?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import pcrlv2_oracle_2d as orc
from pcrlv2_b200.models import PCRLv2
from pcrlv2_b200.models import pcrlv2_model as M

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bsz = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False

def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def unpad2(p, c=None):
    t = p[:, 0, 1:].permute(0, 3, 1, 2).float().contiguous()
    return t if c is None else t[:, :c]

m = PCRLv2(precision=prec); sd0 = orc.init_state(0); m.load_state_dict(orc.clone_state(sd0)); m = m.cuda().train()
sd = orc.clone_state(sd0)
x1 = orc.synthetic_batch(bsz, seed=42, size=(size, size), local=(32, 32))[0]
with torch.no_grad():
    e = "model.encoder"; enc = m.model.encoder
    o = F.relu(orc._bn(F.conv2d(x1, sd[f"{e}.conv1.weight"], None, 2, 3), sd, f"{e}.bn1", True))
    h = m._cb(x1.cuda().float().contiguous(), enc.conv1, enc.bn1, 7, 2, 3, image=True)
    print(f"stem conv+bn+relu  {rl2(unpad2(h), o):.3e}")
    o = F.max_pool2d(o, 3, 2, 1); h = M._MaxPoolFn.apply(h)
    print(f"maxpool            {rl2(unpad2(h), o):.3e}")
    for name, _ci, _co, stride in orc.LAYERS:
        for b in range(2):
            o = orc._basic_block(o, sd, f"{e}.{name}.{b}", stride if b == 0 else 1, True)
            h = m._block(h, getattr(enc, name)[b])
            print(f"{name}.{b}  {tuple(o.shape)}  {rl2(unpad2(h), o):.3e}")
    for i, blk in enumerate(m.model.decoder.blocks):
        o, opro, opre, omask = orc.decoder_block(o, sd, f"model.decoder.blocks.{i}", True)
        h, pro, pre, mask = m._decode_block(h, blk, i, True)
        om = F.interpolate(omask, scale_factor=2 ** (4 - i), mode="bilinear")
        print(f"decoder block {i} {tuple(o.shape)} x {rl2(unpad2(h, o.shape[1]), o):.3e} pro {rl2(pro, opro):.3e} pre {rl2(pre, opre):.3e} mask {rl2(mask, om):.3e}")
    seg = m.model.segmentation_head[0]
    mk = M._ConvC3Fn.apply(h, seg.weight, seg.bias, 16)
    om = F.conv2d(o, sd["model.segmentation_head.0.weight"], sd["model.segmentation_head.0.bias"], 1, 1)
    print(f"segmentation head  {rl2(mk, om):.3e}")
