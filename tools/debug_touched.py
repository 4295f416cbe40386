import sys, os, random, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcrlv2_oracle as orc
from pcrlv2_b200.models import PCRLv23d
from pcrlv2_b200.models import pcrlv2_model_3d as M
from pcrlv2_b200 import train_3d as T

orig = M._LUConvFn.backward
def dbg(ctx, *gs):
    if ctx.cfg.tail:
        print("tail backward cout", ctx.dims[-1], "final", ctx.cfg.final, [None if g is None else tuple(g.shape) for g in gs])
    return orig(ctx, *gs)
M._LUConvFn.backward = staticmethod(dbg)
m = PCRLv23d(); m.load_state_dict(orc.init_state(0)); m = m.cuda().train()
opt = T.FlatSGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
random.seed(1234)
for seed in (42, 43):
    x1, x2, gt, lv = orc.synthetic_batch(2, seed=seed, vol=(32, 32, 16))
    loss, *_ = T.pcrlv2_step_loss(m, x1.cuda(), x2.cuda(), gt.cuda(), [v.cuda() for v in lv], 0, torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    opt.zero_grad(); loss.backward()
    print("touched ds/final:", [n for n in opt.touched_names(m) if "deep_supervision" in n or "out_tr" in n])
    opt.step()
