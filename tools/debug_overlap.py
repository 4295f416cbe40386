"""Compares the flat gradient buffer after ONE backward with / without side-stream weight gradients."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcrlv2_oracle as orc
from pcrlv2_b200.models import PCRLv23d
from pcrlv2_b200 import train_3d as T

x1, x2, gt, lv = orc.synthetic_batch(2, seed=9, vol=(32, 32, 16))
crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
res = {}
for mode in ("1", "0", "0b"):
    os.environ["PCRL_OVERLAP_WGRAD"] = mode[0]
    sd = orc.init_state(0)
    m = PCRLv23d(); m.load_state_dict(orc.clone_state(sd)); m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    random.seed(100)
    loss, _, _, _ = T.pcrlv2_step_loss(m, x1.cuda(), x2.cuda(), gt.cuda(), [v.cuda() for v in lv], 0, crit, cos)
    opt.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    res[mode] = (opt._flat_g.clone(), list(opt._touched), loss.item(), opt, m)
for a, b in (("1", "0"), ("0", "0b")):
    ga, gb = res[a][0], res[b][0]
    print(a, b, "loss", res[a][2], res[b][2], "touched equal", res[a][1] == res[b][1],
          "rel-L2 of flat grads", ((ga - gb).norm() / gb.norm()).item())
    opt, m = res[b][3], res[b][4]
    names = [n for n, _ in m.named_parameters()]
    for i, n in enumerate(names):
        o0, o1 = opt._offs[i], opt._offs[i + 1]
        d = (ga[o0:o1] - gb[o0:o1]).norm().item()
        r = gb[o0:o1].norm().item()
        if r > 0 and d / r > 1e-3:
            print(f"   {n:50s} rel diff {d / r:.3e}  (norm {r:.3e}) touched {res[a][1][i]} {res[b][1][i]}")
