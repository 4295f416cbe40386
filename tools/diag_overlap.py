"""diagnostic: overlapped vs in-line weight-gradient paths, repeated, optionally after a graph capture"""
import os, sys, random, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcrlv2_oracle as orc
from pcrlv2_b200.models import PCRLv23d
from pcrlv2_b200 import train_3d as T
x1, _, gt, _ = orc.synthetic_batch(2, seed=9, vol=(32, 32, 16))
def run(mode):
    os.environ["PCRL_OVERLAP_WGRAD"] = mode
    sd = orc.init_state(0)
    m = PCRLv23d(precision="bf16"); m.load_state_dict(orc.clone_state(sd)); m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    out, _, masks = m(x1.cuda())
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[1], gt.cuda())
    opt.zero_grad(); loss.backward()
    g = opt._flat_g.clone()
    torch.cuda.synchronize()
    g_after = opt._flat_g.clone()
    return g, g_after, {n: p.grad.detach().clone() for n, p in m.named_parameters()}
def rl2(a, b): return ((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item()
if len(sys.argv) > 1 and sys.argv[1] == "graph":
    b = orc.synthetic_batch(4, seed=1, vol=(32, 32, 16))
    sd = orc.init_state(0)
    m = PCRLv23d(precision="fp32"); m.load_state_dict(orc.clone_state(sd)); m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    args = types.SimpleNamespace(lr=1e-2, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    random.seed(1)
    T.train_pcrlv2_inner(args, 0, [(b[0], b[1], b[2], b[2], b[3])], m, opt, torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    print("graph captured:", bool(opt.__dict__.get("_graphed")))
base = run("0")
for i in range(6):
    ov = run("1"); inl = run("0")
    print(f"iter {i}: overlap-vs-inline {rl2(ov[0], base[0]):.3e}  (clone right after backward vs after sync: {rl2(ov[0], ov[1]):.3e})  inline-vs-inline {rl2(inl[0], base[0]):.3e}")
    worst = sorted(((rl2(ov[2][n], base[2][n]), n) for n in base[2] if base[2][n].abs().max() > 0), reverse=True)[:3]
    print("    worst params:", [(f"{e:.2e}", n) for e, n in worst])
