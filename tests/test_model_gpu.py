"""Whole-path parity of the CUDA implementation against the CPU oracle (and the committed golden
fixtures produced from the reference itself, oracle/make_golden.py).

Tolerances (stated, bf16 activation storage + bf16 tensor-core operands, fp32 accumulation):
  * output mask / deep-supervision masks / loss terms: relative L2 <= 3e-2 -- the reference's own
    fp32 -> autocast(bf16) drift on these tensors is 8.5e-3 .. 9e-3;
  * projection / prediction features: relative L2 <= 0.25 at batch 4 -- BatchNorm1d over a tiny
    batch amplifies rounding noise (A.3 measures 8e-2 .. 1.1e-1 for the reference's own bf16 run);
  * parameter gradients: per tensor no further from the fp32 oracle than 1.35x what bf16 storage
    costs the reference's own arithmetic (oracle/bf16_emulation.py) + 0.03; the inherent bf16
    drift (amplified by BatchNorm backward) reaches 0.45 at the stem, stated bound 0.6.
Every comparison is also written to gpurun_out/model_parity.txt for inspection.
"""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LOG = os.path.join(ROOT, "gpurun_out", "model_parity.txt")

from oracle import pcrlv2_oracle as orc  # noqa: E402  (tests may import the oracle)

if torch.cuda.is_available():
    from pcrlv2_b200.models import PCRLv23d
    from pcrlv2_b200 import train_3d as T


def log(msg):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as f:
        f.write(msg + "\n")
    print(msg)


def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(norm="bn", act="relu", seed=0, precision="bf16"):
    sd = orc.init_state(seed, norm=norm, act=act)
    m = PCRLv23d(norm=norm, act=act, precision=precision)
    m.load_state_dict(orc.clone_state(sd))
    return m.cuda().train(), sd


@pytest.mark.parametrize("norm", ["bn", "in"])
def test_forward_vs_oracle_and_golden(norm):
    m, sd0 = build(norm)
    x1, x2, gt, lv = orc.synthetic_batch(2, seed=42)
    sd = orc.clone_state(sd0)
    with torch.no_grad():
        out, feats, masks = m(x1.cuda())
        o_out, o_feats, o_masks = orc.forward(sd, x1, False, True, "relu", norm)
        lout, lfeats, lmasks = m(torch.cat(lv, 0).cuda(), local=True)
        o_lout, o_lfeats, _ = orc.forward(sd, torch.cat(lv, 0), True, True, "relu", norm)
    assert lmasks == []
    errs = {"out": rl2(out, o_out), "local_out": rl2(lout, o_lout)}
    for s in range(3):
        errs[f"mask{s}"] = rl2(masks[s], o_masks[s])
    log(f"[forward {norm}] " + " ".join(f"{k}={v:.3e}" for k, v in errs.items()))
    for s in range(3):
        log(f"[forward {norm}] local pro{s}={rl2(lfeats[s][0], o_lfeats[s][0]):.3e} "
            f"pre{s}={rl2(lfeats[s][1], o_lfeats[s][1]):.3e}  (12 rows)")
    assert max(errs.values()) < 3e-2, errs
    for s in range(3):
        assert rl2(lfeats[s][0], o_lfeats[s][0]) < 0.25
        assert rl2(lfeats[s][1], o_lfeats[s][1]) < 0.25
    # golden digests written from the reference itself
    g = np.load(os.path.join(GOLD, "forward_b2.npz"))
    dig = g[f"{norm}.out"]
    f = out.detach().double().cpu().flatten()
    stride = max(1, f.numel() // 256)
    samp = f[::stride][:256].numpy()
    ref = dig[4:]
    err = np.linalg.norm(samp - ref) / np.linalg.norm(ref)
    log(f"[forward {norm}] golden out samples rel-L2 {err:.3e}; sum {f.sum().item():.4f} vs {dig[0]:.4f}")
    assert err < 3e-2
    assert abs(f.sum().item() - dig[0]) / abs(dig[0]) < 1e-2
    if norm == "bn":
        # BatchNorm running statistics after these two forwards
        msd = m.state_dict()
        worst = 0.0
        for k, v in msd.items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                e = (v.double().cpu() - sd[k].double()).abs().max().item() / max(sd[k].abs().max().item(), 1e-3)
                worst = max(worst, e)
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(sd[k]), k
        log(f"[forward bn] worst running-stat rel err {worst:.3e}")
        assert worst < 5e-2


def test_features_batch4():
    m, sd0 = build("bn")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 1, 32, 32, 16, generator=g)
    sd = orc.clone_state(sd0)
    with torch.no_grad():
        _, feats, _ = m(x.cuda())
        _, o_feats, _ = orc.forward(sd, x, False, True)
    for s in range(3):
        e0, e1 = rl2(feats[s][0], o_feats[s][0]), rl2(feats[s][1], o_feats[s][1])
        log(f"[features b4] scale {s}: pro {e0:.3e} pre {e1:.3e}")
        assert e0 < 0.25 and e1 < 0.25


def _grad_table(m, ograds, tag):
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in ograds.values() if g is not None)).item()
    worst = 0.0
    rows = []
    for name, p in m.named_parameters():
        og = ograds[name]
        if og is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, f"{name}: oracle grad None"
            continue
        assert p.grad is not None, name
        e = rl2(p.grad, og)
        share = og.double().norm().item() / total
        rows.append((name, e, share))
        if share >= 1e-4:
            worst = max(worst, e)
    for name, e, share in rows:
        log(f"[{tag}] {name:55s} rel-L2 {e:.3e}  share {share:.2e}")
    return worst


def test_restoration_gradients_vs_oracle():
    """Gradients of the restoration terms (MSE on the output mask + one deep-supervision mask):
    exercises every trunk backward kernel without BatchNorm1d's small-batch amplification.

    The deviation from the fp32 oracle is dominated by the inherent cost of bf16 storage, which
    BatchNorm-backward cancellation amplifies layer by layer towards the stem: the reference's own
    arithmetic with bf16 storage points (oracle/bf16_emulation.py) sits 0.45 from fp32 at
    down_tr64.ops.0.  Asserted per tensor: err(CUDA, fp32) <= 1.35 * err(emulation, fp32) + 0.03,
    i.e. the kernels add nothing beyond the storage format, and an absolute bound of 0.6.  (CUDA
    and the emulation are two different roundings of a chaotic map, so they do not track each
    other more tightly than either tracks fp32.)"""
    from oracle import bf16_emulation as emu
    m, sd0 = build("bn")
    x1, _, gt, _ = orc.synthetic_batch(2, seed=42)
    grads = {}
    for tag, fwd in (("fp32", lambda sd: orc.forward(sd, x1, False, True)), ("emu", lambda sd: emu.forward(sd, x1))):
        sd = orc.clone_state(sd0)
        keys = [k for k in sd if orc.is_param(k)]
        for k in keys:
            sd[k].requires_grad_(True)
        o_out, _, o_masks = fwd(sd)
        o_loss = torch.nn.functional.mse_loss(o_out, gt) + torch.nn.functional.mse_loss(o_masks[0], gt)
        grads[tag] = dict(zip(keys, torch.autograd.grad(o_loss, [sd[k] for k in keys], allow_unused=True)))
        grads[tag + "_loss"] = o_loss.item()
    out, _, masks = m(x1.cuda())
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[0], gt.cuda())
    loss.backward()
    log(f"[mse-grad] loss {loss.item():.6f} vs fp32 {grads['fp32_loss']:.6f} / emulation {grads['emu_loss']:.6f}")
    for k in [k for k in grads["emu"] if k.endswith("conv1.bias") and "deep_supervision" not in k]:
        grads["emu"][k] = torch.zeros_like(sd0[k])      # the emulation omits the cancelling bias
    _grad_table(m, grads["emu"], "mse-grad vs bf16-emulation")
    _grad_table(m, grads["fp32"], "mse-grad vs fp32")
    # per tensor: the CUDA path may not sit further from fp32 than bf16 storage inherently costs
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads["fp32"].values() if g is not None)).item()
    worst_ratio, worst_abs = 0.0, 0.0
    for name, p in m.named_parameters():
        gf, ge = grads["fp32"][name], grads["emu"][name]
        if gf is None or gf.double().norm().item() / total < 1e-4:
            continue
        e_cuda, e_emu = rl2(p.grad, gf), rl2(ge, gf)
        worst_abs = max(worst_abs, e_cuda)
        worst_ratio = max(worst_ratio, e_cuda / (e_emu + 0.03))
        log(f"[mse-grad] {name:50s} cuda-vs-fp32 {e_cuda:.3e}  emulation-vs-fp32 {e_emu:.3e}")
        assert e_cuda <= 1.35 * e_emu + 0.03, (name, e_cuda, e_emu)
    log(f"[mse-grad] worst cuda-vs-fp32 {worst_abs:.3e}; worst ratio to the bf16-storage floor {worst_ratio:.2f}")
    assert worst_abs < 0.6


def test_step_gradients_vs_oracle():
    """Full step loss at batch 4 (32x32x16 volumes, 16^3 local views)."""
    m, sd0 = build("bn")
    x1, x2, gt, lv = orc.synthetic_batch(4, seed=42, vol=(32, 32, 16))
    sd = orc.clone_state(sd0)
    rng = random.Random(1234)
    scal, draws, ograds = orc.train_step(sd, {}, x1, x2, gt, lv, 0, 0.0, rng)
    random.seed(1234)
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    loss, loss1, loss2, local_loss = T.pcrlv2_step_loss(
        m, x1.cuda(), x2.cuda(), gt.cuda(), [v.cuda() for v in lv], 0, crit, cos)
    loss.backward()
    log(f"[step] loss {loss.item():.6f} vs {scal['loss']:.6f}; loss1 {loss1.item():.6f} vs {scal['loss1']:.6f}; "
        f"loss2 {loss2.item():.6f} vs {scal['loss2']:.6f}; local {float(local_loss):.6f} vs {scal['local_loss']:.6f}")
    assert abs(loss1.item() - scal["loss1"]) < 3e-2 * abs(scal["loss1"])
    assert abs(loss.item() - scal["loss"]) < 5e-2
    worst = _grad_table(m, ograds, "grad")
    log(f"[step] worst significant gradient rel-L2 {worst:.3e}")
    assert worst < 0.8   # bf16 storage drift amplified by BN backward + BatchNorm1d(batch 4); see the test above


def test_two_step_trajectory_vs_golden():
    """FlatSGD + the trainer loop against the REAL reference trainer's result (golden fixture)."""
    m, _ = build("bn")
    g = np.load(os.path.join(GOLD, "train_2steps_b2.npz"))
    batches = [orc.synthetic_batch(2, seed=42), orc.synthetic_batch(2, seed=43)]
    loader = [(b[0], b[1], b[2], b[2], b[3]) for b in batches]
    import types
    args = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    opt = T.FlatSGD(m.parameters(), lr=args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    random.seed(1234)
    mg, local = T.train_pcrlv2_inner(args, 0, loader, m, opt, torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    log(f"[traj] mg_avg {mg:.6f} vs {float(g['mg_avg']):.6f}; local_avg {float(local):.6f} vs {float(g['local_avg']):.6f}")
    assert abs(mg - float(g["mg_avg"])) < 3e-2 * abs(float(g["mg_avg"]))
    # SURVEY note N3: the set of parameters that received a gradient (and hence own a momentum
    # buffer) must equal the reference trainer's; the others must be bit-identical to their
    # initial value (no weight decay, no momentum decay).
    moved = {k[4:] for k in g.files if k.startswith("mom.")}
    names = {id(p): n for n, p in m.named_parameters()}
    has_buf = {names[id(p)] for p in opt.state if "momentum_buffer" in opt.state[p]}
    log(f"[traj] touched sets differ by: {sorted(moved ^ has_buf)}")
    assert moved == has_buf, sorted(moved ^ has_buf)
    sd = m.state_dict()
    init = orc.init_state(0)
    for k, v in sd.items():
        if orc.is_param(k) and k not in moved:
            assert torch.equal(v.detach().cpu(), init[k]), f"{k} must not move"
    # The per-parameter update direction is NOT asserted at batch 2: BatchNorm1d over two rows
    # makes the contrastive gradients ill-conditioned (outputs are +-1, the Jacobian scales with
    # eps/(var+eps)^1.5), so any reduced-precision run decorrelates from fp32 there.  Logged only.
    for k, v in sd.items():
        if not orc.is_param(k) or k not in moved:
            continue
        dig = g[f"state.{k}"]
        f = v.detach().double().cpu().flatten()
        stride = max(1, f.numel() // 256)
        samp = f[::stride][:256].numpy()
        ref = dig[4:] if dig.size > 1 else dig
        i0 = init[k].double().flatten()[::stride][:256].numpy()
        du_ref, du = ref - i0, samp - i0
        e = np.linalg.norm(du - du_ref) / max(np.linalg.norm(du_ref), 1e-30)
        log(f"[traj] {k:55s} update rel-L2 {e:.3e} |du_ref| {np.linalg.norm(du_ref):.2e}")


# ------------------------------------------------------------------------------------------------
# precision="fp32": fp32 activation storage, TF32 tensor-core operands (what the reference itself
# computes with on an Ampere-or-newer GPU: torch's default cudnn.allow_tf32 = True).
def test_fp32_forward_vs_oracle():
    """TF32 operand rounding (2^-11) sets the floor: the output mask lands at 1.0e-3 of the fp32
    oracle, the deep-supervision masks at 0.9e-3 .. 2.4e-3 (the reference's own TF32 drift on
    these tensors is 8e-4).  Stated bound 3e-3; north_star's 1e-3 target is
    met for the two coarse masks and missed by 3 % on `out`."""
    m, sd0 = build("bn", precision="fp32")
    x1, _, gt, lv = orc.synthetic_batch(2, seed=42)
    sd = orc.clone_state(sd0)
    with torch.no_grad():
        out, feats, masks = m(x1.cuda())
        o_out, o_feats, o_masks = orc.forward(sd, x1, False, True)
    errs = {"out": rl2(out, o_out)}
    for s in range(3):
        errs[f"mask{s}"] = rl2(masks[s], o_masks[s])
    log("[fp32 forward] " + " ".join(f"{k}={v:.3e}" for k, v in errs.items()))
    assert max(errs.values()) < 3e-3, errs
    assert errs["out"] < 1.5e-3
    g = np.load(os.path.join(GOLD, "forward_b2.npz"))
    dig = g["bn.out"]
    f = out.detach().double().cpu().flatten()
    samp = f[::max(1, f.numel() // 256)][:256].numpy()
    err = np.linalg.norm(samp - dig[4:]) / np.linalg.norm(dig[4:])
    log(f"[fp32 forward] golden out samples rel-L2 {err:.3e}")
    assert err < 1.5e-3
    worst = 0.0
    for k, v in m.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            worst = max(worst, (v.double().cpu() - sd[k].double()).abs().max().item() / max(sd[k].abs().max().item(), 1e-3))
    log(f"[fp32 forward] worst running-stat rel err {worst:.3e}")
    assert worst < 2e-3


def test_fp32_restoration_gradients():
    """Same test as the bf16 one with TF32 operands.  The floor is what TF32 operand rounding costs
    the reference's own arithmetic (oracle/tf32_emulation.py: 0.12 at the stem, where the CUDA path
    measures 0.13; BatchNorm-backward cancellation amplifies the 2^-11 rounding layer by layer).
    Asserted per tensor: err(CUDA, fp32) <= 1.35 * err(emulation, fp32) + 0.01, absolute bound 0.2."""
    from oracle import tf32_emulation as emu
    m, sd0 = build("bn", precision="fp32")
    x1, _, gt, _ = orc.synthetic_batch(2, seed=42)
    grads = {}
    for tag, fwd in (("fp32", lambda sd: orc.forward(sd, x1, False, True)), ("emu", lambda sd: emu.forward(sd, x1))):
        sd = orc.clone_state(sd0)
        keys = [k for k in sd if orc.is_param(k)]
        for k in keys:
            sd[k].requires_grad_(True)
        o_out, _, o_masks = fwd(sd)
        o_loss = torch.nn.functional.mse_loss(o_out, gt) + torch.nn.functional.mse_loss(o_masks[0], gt)
        grads[tag] = dict(zip(keys, torch.autograd.grad(o_loss, [sd[k] for k in keys], allow_unused=True)))
        grads[tag + "_loss"] = o_loss.item()
    og = grads["fp32"]
    out, _, masks = m(x1.cuda())
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[0], gt.cuda())
    loss.backward()
    log(f"[fp32 mse-grad] loss {loss.item():.7f} vs {grads['fp32_loss']:.7f} (tf32 emulation {grads['emu_loss']:.7f})")
    assert abs(loss.item() - grads["fp32_loss"]) < 1e-5
    worst = _grad_table(m, og, "fp32 mse-grad")
    log(f"[fp32 mse-grad] worst significant gradient rel-L2 {worst:.3e}")
    assert worst < 0.2   # 0.13 at the stem (bf16 storage: 0.45)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in og.values() if g is not None)).item()
    worst_ratio = 0.0
    for name, p in m.named_parameters():
        gf, ge = og[name], grads["emu"][name]
        if gf is None or gf.double().norm().item() / total < 1e-4:
            continue
        e_cuda, e_emu = rl2(p.grad, gf), rl2(ge, gf)
        worst_ratio = max(worst_ratio, e_cuda / (e_emu + 0.01))
        log(f"[fp32 mse-grad] {name:50s} cuda-vs-fp32 {e_cuda:.3e}  tf32-emulation-vs-fp32 {e_emu:.3e}")
        assert e_cuda <= 1.35 * e_emu + 0.01, (name, e_cuda, e_emu)
    log(f"[fp32 mse-grad] worst ratio to the TF32-operand floor {worst_ratio:.2f}")


# BASELINE configs[3]: 128x128x64 crops (the large-volume regime; local views 32^3 keep the 1/32
# voxel ratio).  Parity-test case, not a bench line.
_LARGE = {}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_large_volume_128x128x64(precision):
    m, sd0 = build("bn", precision=precision)
    x1, _, gt, lv = orc.synthetic_batch(2, seed=11, vol=(128, 128, 64), local=(32, 32, 32), n_local=1)
    if "ref" not in _LARGE:          # the CPU oracle at this size takes ~30 s: once for both precisions
        sd = orc.clone_state(sd0)
        keys = [k for k in sd if orc.is_param(k)]
        for k in keys:
            sd[k].requires_grad_(True)
        o_out, _, o_masks = orc.forward(sd, x1, False, True)
        o_loss = torch.nn.functional.mse_loss(o_out, gt) + torch.nn.functional.mse_loss(o_masks[1], gt)
        og = dict(zip(keys, torch.autograd.grad(o_loss, [sd[k] for k in keys], allow_unused=True)))
        _LARGE["ref"] = (o_out.detach(), [t.detach() for t in o_masks], o_loss.detach(), og)
    o_out, o_masks, o_loss, og = _LARGE["ref"]
    out, feats, masks = m(x1.cuda())
    errs = {"out": rl2(out, o_out)}
    for s_ in range(3):
        errs[f"mask{s_}"] = rl2(masks[s_], o_masks[s_])
    log(f"[128x128x64 {precision}] " + " ".join(f"{k}={v:.3e}" for k, v in errs.items()))
    # the 1-channel deep-supervision masks sit at 2.4e-2 already at 64x64x32 (bf16 storage of the
    # 64-channel activation the head reads); 3.2e-2 here
    tol = 5e-2 if precision == "bf16" else 4e-3
    assert max(errs.values()) < tol, errs
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[1], gt.cuda())
    loss.backward()
    log(f"[128x128x64 {precision}] loss {loss.item():.6f} vs {o_loss.item():.6f}")
    assert abs(loss.item() - o_loss.item()) < (2e-3 if precision == "bf16" else 2e-5)
    # the decoder's directly supervised tensors (well conditioned at any size)
    for name in ("out_tr.final_conv.weight", "up_tr64.ops.1.conv1.weight", "up_tr128.deep_supervision_head.conv1.weight"):
        e = rl2(dict(m.named_parameters())[name].grad, og[name])
        log(f"[128x128x64 {precision}] {name} grad rel-L2 {e:.3e}")
        assert e < (0.1 if precision == "bf16" else 2e-2), (name, e)
    # local views of 32^3 through the local path
    with torch.no_grad():
        lout, lfeats, lmasks = m(lv[0].cuda(), local=True)
        o_lout, o_lfeats, _ = orc.forward(orc.clone_state(sd0), lv[0], True, True)
    assert lmasks == []
    e = rl2(lout, o_lout)
    log(f"[128x128x64 {precision}] local 32^3 out rel-L2 {e:.3e}")
    assert e < tol


def test_side_stream_weight_gradients_match(monkeypatch):
    """With a FlatSGD attached, the 3x3x3 weight gradients are accumulated into the flat buffer from a
    side stream (models/pcrlv2_model_3d.py:_wgrad_overlapped).  Same gradients, same set of reached
    parameters, same update as the in-line path (PCRL_OVERLAP_WGRAD=0).  The restoration terms are
    used (well conditioned at batch 2; the contrastive terms are chaotic there: BatchNorm1d over two
    rows), and the run-to-run noise of the in-line path (atomics) calibrates the comparison."""
    # batch 8, fp32 storage: at batch 2 in bf16 the run-to-run noise of ONE path through the BatchNorm
    # backward is heavy-tailed (4e-3 .. 9e-2 rel-L2 over repeated runs, tools/diag_overlap.py), which made a
    # one-sample noise calibration flaky
    x1, _, gt, _ = orc.synthetic_batch(8, seed=9, vol=(32, 32, 16))
    runs = []
    for mode in ("1", "0", "0", "0"):
        monkeypatch.setenv("PCRL_OVERLAP_WGRAD", mode)
        m, _ = build("bn", precision="fp32")
        opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
        out, _, masks = m(x1.cuda())
        loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[1], gt.cuda())
        opt.zero_grad()
        loss.backward()
        g = opt._flat_g.clone()                # reads p.grad right after backward(): must be complete
        touched = list(opt._touched)
        opt.step()
        runs.append((g, touched, opt._flat_p.clone(), loss.item()))
    (g1, t1, p1, l1), (g0, t0, p0, l0), (gb, tb, pb, lb), (gc, _, pc, _) = runs
    assert t1 == t0 == tb
    # fp64 / fp32 atomics make two runs of the SAME path differ (loss ~3e-6, gradients ~4e-3 rel-L2
    # at batch 2 through the BatchNorm backward): the bounds sit well above that noise and far below
    # what a lost or doubled weight gradient would cause (rel-L2 ~ 1)
    assert abs(l1 - l0) < 1e-4
    noise = max(rl2(gb, g0), rl2(gc, g0), rl2(gc, gb))
    diff = rl2(g1, g0)
    log(f"[side-stream wgrad] flat gradient rel-L2 overlapped vs in-line {diff:.3e}; in-line run-to-run {noise:.3e}")
    assert diff < max(5 * noise, 3e-2)
    assert rl2(p1, p0) < max(5 * max(rl2(pb, p0), rl2(pc, p0)), 1e-5)


def test_checkpoint_round_trip_with_torch_sgd():
    """the checkpoint the trainer writes (reference train_3d.py:71-80) is
    interchangeable with the reference's: FlatSGD.state_dict() loads into torch.optim.SGD and steps
    identically, torch.optim.SGD.state_dict() loads into FlatSGD, model.state_dict() round-trips."""
    import io
    x1, _, gt, _ = orc.synthetic_batch(2, seed=4, vol=(32, 32, 16))
    m, _ = build("bn")
    opt = T.FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)

    def backward_once():
        out, _, masks = m(x1.cuda())
        loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[2], gt.cuda())
        opt.zero_grad()
        loss.backward()

    backward_once()
    opt.step()
    # the checkpoint dict of the trainer, through torch.save / torch.load
    buf = io.BytesIO()
    torch.save({"state_dict": m.state_dict(), "optimizer": opt.state_dict(), "epoch": 0}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    assert list(ck["state_dict"].keys()) == [k for k, _, _ in orc.state_spec()]
    # reference side: plain parameters + torch.optim.SGD restored from OUR optimizer state
    names = [n for n, _ in m.named_parameters()]
    ref_params = [torch.nn.Parameter(ck["state_dict"][n].detach().clone().cuda()) for n in names]
    ref_opt = torch.optim.SGD(ref_params, lr=1e-2, momentum=0.9, weight_decay=1e-4)
    ref_opt.load_state_dict(ck["optimizer"])
    n_buf = sum(1 for p in ref_params if "momentum_buffer" in ref_opt.state.get(p, {}))
    assert n_buf == sum(opt._has_buf) and 0 < n_buf < len(ref_params)   # unreached heads have no buffer (N3)
    # a second step with the same gradients on both sides
    backward_once()
    for p, q, touched in zip(m.parameters(), ref_params, opt._touched):
        q.grad = p.grad.detach().clone() if touched else None
    opt.step()
    ref_opt.step()
    worst = max(rl2(p, q) for p, q in zip(m.parameters(), ref_params))
    log(f"[checkpoint] FlatSGD vs torch.optim.SGD restored from its state_dict: worst parameter rel-L2 {worst:.3e}")
    assert worst < 1e-6
    # and back: torch.optim.SGD's state into a fresh FlatSGD
    m2, _ = build("bn", seed=1)
    opt2 = T.FlatSGD(m2.parameters(), lr=1e-3, momentum=0.0, weight_decay=0.0)
    m2.load_state_dict({**m.state_dict()})
    opt2.load_state_dict(ref_opt.state_dict())
    assert opt2.param_groups[0]["lr"] == 1e-2 and opt2.param_groups[0]["momentum"] == 0.9
    assert opt2._has_buf == opt._has_buf
    assert rl2(opt2._flat_m, opt._flat_m) < 1e-6
    assert rl2(opt2._flat_p, opt._flat_p) < 1e-7
    with pytest.raises(ValueError):
        bad = ref_opt.state_dict()
        bad["param_groups"][0]["nesterov"] = True
        opt2.load_state_dict(bad)


@pytest.mark.parametrize("act", ["prelu", "elu"])
def test_activation_variants_vs_reference_fixture(act):
    """act='prelu' / 'elu' (reference models/pcrlv2_model_3d.py:20-27) through the CUDA path, fp32
    mode, against the digests written from the reference model itself (tests/golden/acts_b2.npz)."""
    g = np.load(os.path.join(GOLD, "acts_b2.npz"))
    m, _ = build("bn", act=act, precision="fp32")
    x1, _, gt, _ = orc.synthetic_batch(2, seed=7, vol=(32, 32, 16))
    out, _, masks = m(x1.cuda())
    loss = torch.nn.functional.mse_loss(out, gt.cuda()) + torch.nn.functional.mse_loss(masks[1], gt.cuda())
    loss.backward()
    assert abs(loss.item() - float(g[f"{act}.loss"])) < 2e-5
    f = out.detach().double().cpu().flatten()
    stride = max(1, f.numel() // 256)
    ref = g[f"{act}.out"]
    err = np.linalg.norm(f[::stride][:256].numpy() - ref[4:]) / np.linalg.norm(ref[4:])
    log(f"[act {act}] loss {loss.item():.7f} vs reference {float(g[f'{act}.loss']):.7f}; out samples rel-L2 {err:.3e}")
    assert err < 4e-3
    assert abs(f.sum().item() - ref[0]) / abs(ref[0]) < 1e-3
    # the decoder's last convolution and (prelu) its slope gradient: digest = [sum, |.|sum, sq sum, n, samples]
    params = dict(m.named_parameters())
    for name in ("up_tr64.ops.1.conv1.weight",) + (("up_tr64.ops.1.activation.weight",) if act == "prelu" else ()):
        d = g[f"{act}.grad.{name}"]
        gr = params[name].grad.detach().double().cpu().flatten()
        st = max(1, gr.numel() // 256)
        e = np.linalg.norm(gr[::st][:256].numpy() - d[4:]) / np.linalg.norm(d[4:])
        log(f"[act {act}] {name} grad samples rel-L2 {e:.3e}")
        assert e < 5e-2, (name, e)
