"""diagnostic: which parameters differ between the overlapped and in-line weight-gradient paths"""
import os, sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcrlv2_oracle as orc
from pcrlv2_b200.models import PCRLv23d
from pcrlv2_b200 import train_3d as T
x1, _, gt, _ = orc.synthetic_batch(2, seed=9, vol=(32, 32, 16))
res = {}
for tag, mode in (("ov", "1"), ("in", "0"), ("in2", "0"), ("ov2", "1")):
    os.environ["PCRL_OVERLAP_WGRAD"] = mode
    sd = orc.init_state(0)
    m = PCRLv23d(precision="bf16"); m.load_state_dict(orc.clone_state(sd)); m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    out, _, masks = m(x1.cuda())
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[1], gt.cuda())
    opt.zero_grad(); loss.backward()
    torch.cuda.synchronize()
    res[tag] = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
def rl2(a, b): return ((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item()
for n in res["ov"]:
    a, b, c, d = rl2(res["ov"][n], res["in"][n]), rl2(res["in2"][n], res["in"][n]), rl2(res["ov2"][n], res["in"][n]), rl2(res["ov2"][n], res["ov"][n])
    if max(a, c) > 5 * max(b, 1e-3): print(f"{n:50s} ov-vs-in {a:.3e} in-vs-in {b:.3e} ov2-vs-in {c:.3e} ov2-vs-ov {d:.3e}")
print("done")
