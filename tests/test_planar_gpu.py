"""GPU: the 2-D path -- csrc/planar.cu kernels op by op against torch's own fp32 ops, the
whole PCRLv2 model against the CPU oracle, and one full training step against the fixture the REAL reference
trainer (train_2d.train_pcrlv2_inner over the smp restatement) wrote."""
import os
import random
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import pcrlv2_oracle_2d as orc  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "planar_parity.txt")


def log(msg):
    print(msg)
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, "a") as f:
            f.write(msg + "\n")
    except OSError:
        pass


@pytest.fixture(autouse=True)
def _true_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def q(t):
    """values exactly representable in bf16 (hence in tf32): products are exact in both storage modes"""
    return t.to(torch.bfloat16).float()


def pad2(x, dtype):
    """NCHW fp32 -> H-padded NHWC [N,1,H+1,W,C]"""
    n, c, h, w = x.shape
    out = torch.zeros((n, 1, h + 1, w, c), dtype=dtype, device=x.device)
    out[:, 0, 1:] = x.permute(0, 2, 3, 1).to(dtype)
    return out


def unpad2(p):
    return p[:, 0, 1:].permute(0, 3, 1, 2).float().contiguous()


DT = [torch.float32, torch.bfloat16]
CONVS = [  # (cin, cout, k, s, p, h, w, image)
    (3, 64, 7, 2, 3, 64, 64, True),
    (64, 64, 3, 1, 1, 16, 16, False),
    (64, 128, 3, 2, 1, 16, 16, False),
    (64, 128, 1, 2, 0, 16, 16, False),
    (32, 16, 3, 1, 1, 24, 20, False),
    (128, 64, 3, 1, 1, 9, 7, False),
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", CONVS)
def test_conv2d_im2col_gemm_fwd_bwd(cfg, dtype):
    """Conv2d forward, data gradient and weight gradient through im2col / GEMM / col2im against F.conv2d."""
    from pcrlv2_b200 import kernels as K, kernels2d as K2
    cin, cout, k, s, p, h, w, image = cfg
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 3
    x = q(torch.randn((n, cin, h, w), device="cuda", generator=g))
    wt = q(torch.randn((cout, cin, k, k), device="cuda", generator=g) * 0.1)
    ho, wo = K2.out_size(h, k, s, p), K2.out_size(w, k, s, p)
    dy = q(torch.randn((n, cout, ho, wo), device="cuda", generator=g))
    xr, wr = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, s, p)
    yr.backward(dy)
    xin = x if image else pad2(x, dtype)
    cs = cin
    col, ho2, wo2 = K2.im2col2d(xin, k, s, p, dtype, image=image)
    assert (ho2, wo2) == (ho, wo)
    wmat, wtr = K2.pack_conv2d_weights(wt, cs, dtype)
    coutp = wmat.shape[0]
    stats = torch.zeros((1, coutp, 2), dtype=torch.float64, device="cuda")
    y = K2.gemm_nt_stats(col, wmat, stats).view(n, 1, ho + 1, wo, coutp)
    tol = 6e-4 if dtype == torch.float32 else 6e-3      # exact products, fp32 accumulate; the stored output is rounded (tf32 / bf16)
    got = unpad2(y)[:, :cout]
    assert rl2(got, yr) < tol, rl2(got, yr)
    assert float(y[:, 0, 0].abs().max()) == 0.0 and float(unpad2(y)[:, cout:].abs().max() if coutp > cout else 0.0) == 0.0
    # statistics epilogue: column sums of the stored output
    ssum = unpad2(y).double().sum((0, 2, 3))
    assert torch.allclose(stats[0, :, 0], ssum, rtol=1e-6, atol=1e-4)
    dyp = torch.zeros((n, 1, ho + 1, wo, coutp), dtype=dtype, device="cuda")
    dyp[:, 0, 1:, :, :cout] = dy.permute(0, 2, 3, 1).to(dtype)
    dy2d = dyp.view(-1, coutp)
    dw = K2.conv2d_wgrad(dy2d, col, cout, cin, k, cs)
    assert rl2(dw, wr.grad) < 2e-5, rl2(dw, wr.grad)
    if not image:
        dcol = K.gemm_nt(dy2d, wtr, out_fp32=False)
        dx = K2.col2im2d(dcol, n, h, w, cs, k, s, p)
        assert rl2(unpad2(dx), xr.grad) < tol, rl2(unpad2(dx), xr.grad)
        assert float(dx[:, 0, 0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("hw", [(16, 16), (15, 9), (6, 8)])
def test_maxpool_3x3s2_with_ties(hw, dtype):
    from pcrlv2_b200 import kernels2d as K2
    h, w = hw
    g = torch.Generator(device="cuda").manual_seed(2)
    # post-ReLU, coarsely quantised: many exact ties inside the windows
    x = (torch.randn((2, 16, h, w), device="cuda", generator=g).clamp_min(0) * 2).round() / 2
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    dy = q(torch.randn(yr.shape, device="cuda", generator=g))
    yr.backward(dy)
    xp = pad2(x, dtype)
    y = K2.maxpool_fwd(xp)
    assert torch.equal(unpad2(y), yr.detach())
    dx = K2.maxpool_bwd(xp, pad2(dy, dtype))
    assert rl2(unpad2(dx), xr.grad) < (5e-4 if dtype == torch.float32 else 4e-3)
    assert float(dx[:, 0, 0].abs().max()) == 0.0 and float(y[:, 0, 0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DT)
def test_add_relu_and_nearest(dtype):
    from pcrlv2_b200 import kernels2d as K2
    g = torch.Generator(device="cuda").manual_seed(3)
    a = q(torch.randn((2, 32, 6, 10), device="cuda", generator=g))
    b = q(torch.randn((2, 32, 6, 10), device="cuda", generator=g))
    out = K2.add_relu(pad2(a, dtype), pad2(b, dtype), 0)
    ref = F.relu(a + b)
    assert rl2(unpad2(out), ref) < 4e-3
    gg = q(torch.randn_like(a))
    dg = K2.add_relu(out, pad2(gg, dtype), 1)
    assert torch.equal(unpad2(dg), gg * (unpad2(out) > 0))
    s = K2.add_relu(pad2(a, dtype), pad2(b, dtype), 2)
    assert rl2(unpad2(s), a + b) < 4e-3
    up = K2.up_nearest_fwd(pad2(a, dtype))
    assert torch.equal(unpad2(up), F.interpolate(a, scale_factor=2, mode="nearest"))
    assert float(up[:, 0, 0].abs().max()) == 0.0
    gf = q(torch.randn((2, 32, 12, 20), device="cuda", generator=g))
    dx = K2.up_nearest_bwd(pad2(gf, dtype))
    ref = gf.view(2, 32, 6, 2, 10, 2).sum((3, 5))
    assert rl2(unpad2(dx), ref) < (5e-4 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("sf", [1, 2, 4, 16])
def test_bilinear_fwd_bwd(sf):
    from pcrlv2_b200 import kernels2d as K2
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((2, 3, 5, 7), device="cuda", generator=g)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=sf, mode="bilinear")
    dy = torch.randn(yr.shape, device="cuda", generator=g)
    yr.backward(dy)
    y = K2.bilinear_fwd(x, sf)
    assert (y - yr.detach()).abs().max().item() < 2e-6
    dx = K2.bilinear_bwd(dy, sf)
    assert rl2(dx, xr.grad) < 2e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("cfg", [(16, 32, 3), (256, 256, 1), (32, 32, 1), (16, 32, 1)])
def test_conv_to_3_channels(cfg, dtype):
    from pcrlv2_b200 import kernels2d as K2
    c, cs, k = cfg
    g = torch.Generator(device="cuda").manual_seed(5)
    n, h, w = 2, 10, 12
    a = torch.zeros((n, cs, h, w), device="cuda")
    a[:, :c] = q(torch.randn((n, c, h, w), device="cuda", generator=g))
    wt = torch.randn((3, c, k, k), device="cuda", generator=g) * 0.2
    bias = torch.randn((3,), device="cuda", generator=g)
    ar, wr, br = a[:, :c].clone().requires_grad_(True), wt.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yr = F.conv2d(ar, wr, br, 1, k // 2)
    dy = torch.randn(yr.shape, device="cuda", generator=g)
    yr.backward(dy)
    ap = pad2(a, dtype)
    y = K2.conv_c3_fwd(ap, wt, bias, c)
    assert rl2(y, yr) < 2e-6
    da, dw, db = K2.conv_c3_bwd(ap, wt, dy, c)
    assert rl2(unpad2(da)[:, :c], ar.grad) < (5e-4 if dtype == torch.float32 else 4e-3)
    if cs > c:
        assert float(unpad2(da)[:, c:].abs().max()) == 0.0
    assert rl2(dw, wr.grad) < 1e-5 and rl2(db, br.grad) < 1e-5


# ------------------------------------------------------------------------------ whole model
def build2d(precision, seed=0):
    from pcrlv2_b200.models import PCRLv2
    m = PCRLv2(precision=precision)
    sd0 = orc.init_state(seed)
    m.load_state_dict(orc.clone_state(sd0))
    return m.cuda().train(), sd0


EMU = {"fp32": "tf32", "bf16": "bf16"}
# A 27-convolution network without skip connections amplifies operand rounding: the REFERENCE'S OWN arithmetic with
# TF32-rounded operands (what its cuDNN convolutions compute on a GPU by default) lands 2.3e-2 from true fp32 at the
# output mask of this configuration, with bf16 operands 0.18; its gradients 0.19 (median) .. 0.36 / 0.59 .. 1.0
# (oracle/operand_emulation_2d.py).  "Within tolerance of the reference" can therefore only mean: as close to the
# fp32 oracle as the emulation of the same precision mode is.  Kernel correctness proper is asserted op by op above.
SLACK = {"fp32": (1.6, 2e-3, 0.05), "bf16": (2.0, 2e-2, 0.2)}     # (factor on the emulation's error, forward abs, gradient abs)


def test_model2d_forward_fp32x3_within_1e3_of_the_fp32_reference():
    """precision='fp32x3' (3xTF32 split operands, unrounded fp32 storage): the mode in which "outputs within 1e-3
    relative of the reference fp32" is asserted for the 2-D model -- final mask, the five middle masks, global and
    local views, BatchNorm buffers."""
    m, sd0 = build2d("fp32x3")
    sd = orc.clone_state(sd0)
    x1, _x2, _gt, lv = orc.synthetic_batch(4, seed=42, size=(64, 64), local=(32, 32))
    with torch.no_grad():
        o_dec, o_mask, o_mm = orc.forward(sd, x1)
        o_ldec, _, o_lmm = orc.forward(sd, torch.cat(lv, 0), local=True)
        dec, mask, mm = m(x1.cuda())
        ldec, _, lmm = m(torch.cat(lv, 0).cuda(), local=True)
    e = rl2(mask, o_mask)
    log(f"[2d fwd fp32x3] mask rel-L2 vs fp32 oracle {e:.3e}")
    assert e < 1e-3
    for s in range(5):
        em, el = rl2(mm[s], o_mm[s]), rl2(lmm[s], o_lmm[s])
        ep, eq, elp = rl2(dec[s][0], o_dec[s][0]), rl2(dec[s][1], o_dec[s][1]), rl2(ldec[s][0], o_ldec[s][0])
        log(f"[2d fwd fp32x3] scale {s}: middle mask {em:.3e} (local {el:.3e}) pro {ep:.3e} pre {eq:.3e} local pro {elp:.3e}")
        assert em < 1e-3 and el < 1e-3
        assert ep < 5e-3 and elp < 5e-3 and eq < 2e-2          # BatchNorm1d over 4 / 24 nearly identical rows
    worst = 0.0
    for k, v in m.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            worst = max(worst, (v.cpu() - sd[k]).abs().max().item() / max(1.0, sd[k].abs().max().item()))
    log(f"[2d fwd fp32x3] worst BatchNorm buffer deviation {worst:.3e}")
    assert worst < 1e-4


def test_full_step_2d_fp32x3_gradients():
    """The full iteration in the parity mode against the oracle's full tensors: loss terms to 1e-4, every parameter's
    gradient within max(5e-2, 4 x the reference's own fp32-vs-fp64 floor) -- an order of magnitude below the TF32
    mode's 0.1-0.2 (what remains is the tensor core's truncating fp32 accumulation, amplified by this network)."""
    from pcrlv2_b200 import train_2d as T2
    from pcrlv2_b200.train_3d import FlatSGD
    g = np.load(os.path.join(GOLD, "train2d_2steps_b8.npz"))
    lr = float(g["lr"])
    m, sd0 = build2d("fp32x3")
    sd = orc.clone_state(sd0)
    b = orc.synthetic_batch(8, seed=42, size=(64, 64), local=(32, 32))
    scal, draws, ograds = orc.train_step(sd, {}, b[0], b[1], b[2], b[3], 0, lr, random.Random(1234))
    opt = FlatSGD(m.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    random.seed(1234)
    loss, loss1, loss2, local_loss = T2.pcrlv2_step_loss(m, b[0].cuda(), b[1].cuda(), b[2].cuda(),
                                                         [v.cuda() for v in b[3]], 0, crit, cos)
    opt.zero_grad()
    loss.backward()
    got = dict(loss=loss.item(), loss1=loss1.item(), loss2=loss2.item(), local_loss=float(local_loss))
    for k, v in got.items():
        log(f"[2d step fp32x3] {k} {v:.7f} vs oracle {scal[k]:.7f}")
        assert abs(v - scal[k]) < 1e-4, (k, v, scal[k])
    errs, failures = [], []
    for i, (n, p) in enumerate(m.named_parameters()):
        og = ograds[n]
        if og is None:
            assert not opt._touched[i], n
            continue
        if orc.is_cancelling(n):
            continue
        floor = float(g[f"floor1.{n}"])
        eg = rl2(p.grad, og)
        errs.append(eg)
        if eg > max(5e-2, 4 * floor):
            failures.append((n, eg, floor))
    log(f"[2d step fp32x3] {len(errs)} parameters: gradient rel-L2 vs fp32 oracle median {np.median(errs):.3e}, worst {max(errs):.3e}")
    assert not failures, failures


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model2d_forward_vs_oracle(precision):
    """Train-mode forward of the whole model (global 64x64 batch and 24 local 32x32 views) against the oracle:
    final mask, the five upsampled middle masks, (pro, pre) of every scale, BatchNorm buffers."""
    from oracle import operand_emulation_2d as emu
    m, sd0 = build2d(precision)
    sd = orc.clone_state(sd0)
    x1, _x2, _gt, lv = orc.synthetic_batch(4, seed=42, size=(64, 64), local=(32, 32))
    with torch.no_grad():
        o_dec, o_mask, o_mm = orc.forward(sd, x1)
        o_ldec, _, o_lmm = orc.forward(sd, torch.cat(lv, 0), local=True)
        with emu.rounding(EMU[precision]):
            sde = orc.clone_state(sd0)
            e_dec, e_mask, e_mm = orc.forward(sde, x1)
            e_ldec, _, e_lmm = orc.forward(sde, torch.cat(lv, 0), local=True)
        dec, mask, mm = m(x1.cuda())
        ldec, lmask, lmm = m(torch.cat(lv, 0).cuda(), local=True)
    assert lmask is None and len(mm) == 5 and len(lmm) == 5
    f, tol, _ = SLACK[precision]

    def check(name, got, ref, emulated, extra=1.0):
        e, ee = rl2(got, ref), rl2(emulated, ref)
        log(f"[2d fwd {precision}] {name:24s} CUDA vs fp32 oracle {e:.3e}; reference arithmetic with {EMU[precision]} operands {ee:.3e}")
        assert e <= extra * f * ee + tol, (name, e, ee)

    check("mask", mask, o_mask, e_mask)
    for s in range(5):
        check(f"middle mask {s}", mm[s], o_mm[s], e_mm[s])
        check(f"local middle mask {s}", lmm[s], o_lmm[s], e_lmm[s])
        # BatchNorm1d over 4 / 24 nearly identical rows: two noisy realisations, a wider factor
        check(f"pro {s}", dec[s][0], o_dec[s][0], e_dec[s][0], 2.0)
        check(f"pre {s}", dec[s][1], o_dec[s][1], e_dec[s][1], 2.0)
        check(f"local pro {s}", ldec[s][0], o_ldec[s][0], e_ldec[s][0], 2.0)
    worst = 0.0
    for k, v in m.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(sd[k]), k
        elif k.endswith("running_mean") or k.endswith("running_var"):
            worst = max(worst, (v.cpu() - sd[k]).abs().max().item() / max(1.0, sd[k].abs().max().item()))
    log(f"[2d fwd {precision}] worst BatchNorm buffer deviation {worst:.3e}")
    assert worst < (5e-3 if precision == "fp32" else 5e-2)


def test_model2d_need_masks_false_keeps_the_state_trajectory():
    """The trainer's forwards for x2 / local views skip the mask heads' output (need_masks=False, an extension).
    They must leave exactly the state the full forward leaves: every BatchNorm buffer and counter, and the same
    (pro, pre) outputs."""
    x = orc.synthetic_batch(4, seed=7, size=(64, 64), local=(32, 32))[0].cuda()
    outs = {}
    for need in (True, False):
        m, _ = build2d("fp32")
        with torch.no_grad():
            dec, mask, mm = m(x, need_masks=need)
        outs[need] = (dec, {k: v.detach().clone() for k, v in m.state_dict().items()})
        assert (mask is None) == (not need) and len(mm) == (5 if need else 0)
    for (p1, q1), (p2, q2) in zip(outs[True][0], outs[False][0]):
        assert rl2(p2, p1) < 5e-2 and rl2(q2, q1) < 0.2           # run-to-run noise level (tools/diag_2d_det.py)
    for k, v in outs[True][1].items():
        w = outs[False][1][k]
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(w), k
        elif k.endswith("running_mean") or k.endswith("running_var"):
            # two forwards of the same input already differ at this level (run-to-run noise, tools/diag_2d_det.py);
            # the BatchNorm1d statistics are taken over 4 nearly identical rows and move the most (2.5e-3 seen)
            tol = 2e-2 if ("predictor_head" in k or ".bn." in k) else 5e-3
            assert (v - w).abs().max().item() <= tol * max(1.0, v.abs().max().item()), k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_model2d_eval_mode_forward(precision):
    """model.eval(): BatchNorm2d / BatchNorm1d normalise with the running statistics; buffers do not move."""
    from oracle import operand_emulation_2d as emu
    m, sd0 = build2d(precision)
    # non-trivial running statistics: one train-mode forward on both sides first
    x0, x1 = (orc.synthetic_batch(4, seed=s, size=(64, 64), local=(32, 32))[0] for s in (11, 12))
    sd = orc.clone_state(sd0)
    with torch.no_grad():
        orc.forward(sd, x0)
        m(x0.cuda())
        m.eval()
        before = {k: v.detach().clone() for k, v in m.state_dict().items()}
        sde = orc.clone_state(sd)
        o_dec, o_mask, o_mm = orc.forward(sd, x1, training=False)
        with emu.rounding(EMU[precision]):
            e_dec, e_mask, e_mm = orc.forward(sde, x1, training=False)
        dec, mask, mm = m(x1.cuda())
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    f, tol, _ = SLACK[precision]
    e, ee = rl2(mask, o_mask), rl2(e_mask, o_mask)
    log(f"[2d eval {precision}] mask CUDA vs fp32 oracle {e:.3e}; emulation {ee:.3e}")
    # the running statistics of the two sides already differ by the train-mode forward's noise: a wider factor
    assert e <= 2.5 * f * ee + 5 * tol, (e, ee)
    for s_ in range(5):
        em, ep = rl2(mm[s_], o_mm[s_]), rl2(dec[s_][0], o_dec[s_][0])
        log(f"[2d eval {precision}] scale {s_}: middle mask {em:.3e} (emulation {rl2(e_mm[s_], o_mm[s_]):.3e}) pro {ep:.3e}")
        assert em <= 2.5 * f * rl2(e_mm[s_], o_mm[s_]) + 5 * tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_step_2d_vs_reference_fixture(precision):
    """One full iteration (three forwards, four loss terms, backward, SGD) at the fixture configuration (b=8,
    64x64 + 6 x 32x32, lr 1e-2).  The oracle is pinned on the REAL train_2d.train_pcrlv2_inner by the fixture
    (loss terms, draws).  Every parameter's gradient and update is compared with the oracle's full tensors and must
    be as close as the reference's own arithmetic with the same operand precision is (see SLACK); parameters the
    reference leaves without a gradient must not move (note N3)."""
    from oracle import operand_emulation_2d as emu
    from pcrlv2_b200 import train_2d as T2
    from pcrlv2_b200.train_3d import FlatSGD
    g = np.load(os.path.join(GOLD, "train2d_2steps_b8.npz"))
    lr = float(g["lr"])
    f, tol_fwd, tol_grad = SLACK[precision]
    m, sd0 = build2d(precision)
    sd = orc.clone_state(sd0)
    b = orc.synthetic_batch(8, seed=42, size=(64, 64), local=(32, 32))
    scal, draws, ograds = orc.train_step(sd, {}, b[0], b[1], b[2], b[3], 0, lr, random.Random(1234))
    assert list(draws) == list(g["draws"][0])
    for k in ("loss", "loss1", "loss2", "loss4", "local_loss"):
        assert abs(scal[k] - float(g[f"step0.{k}"])) < 2e-5, k
    with emu.rounding(EMU[precision]):
        escal, _, egrads = orc.train_step(orc.clone_state(sd0), {}, b[0], b[1], b[2], b[3], 0, lr, random.Random(1234))
    opt = FlatSGD(m.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    random.seed(1234)
    loss, loss1, loss2, local_loss = T2.pcrlv2_step_loss(m, b[0].cuda(), b[1].cuda(), b[2].cuda(),
                                                         [v.cuda() for v in b[3]], 0, crit, cos)
    opt.zero_grad()
    loss.backward()
    got = dict(loss=loss.item(), loss1=loss1.item(), loss2=loss2.item(), local_loss=float(local_loss))
    for k, v in got.items():
        log(f"[2d step {precision}] {k} {v:.7f} vs oracle {scal[k]:.7f} (emulation {escal[k]:.7f})")
    for k, v in got.items():
        assert abs(v - scal[k]) <= 3 * abs(escal[k] - scal[k]) + tol_fwd, (k, v, scal[k], escal[k])
    grads = {n: (p.grad.detach().clone() if opt._touched[i] else None) for i, (n, p) in enumerate(m.named_parameters())}
    opt.step()
    msd = dict(m.named_parameters())
    ratios, failures = [], []
    for n in msd:
        og = ograds[n]
        if og is None:
            assert grads[n] is None, f"{n}: the reference gives this parameter no gradient"
            assert torch.equal(msd[n].detach().cpu(), sd0[n]), f"{n} must not move"
            continue
        assert grads[n] is not None, n
        if orc.is_cancelling(n):
            continue
        floor = float(g[f"floor1.{n}"])
        eg, ee = rl2(grads[n], og), rl2(egrads[n], og)
        du, du_ref = msd[n].detach().cpu().double() - sd0[n].double(), sd[n].double() - sd0[n].double()
        eu = ((du - du_ref).norm() / du_ref.norm().clamp_min(1e-30)).item()
        log(f"[2d step {precision}] {n:60s} grad rel-L2 {eg:.3e} (emulation {ee:.3e}, fp32 floor {floor:.1e}) update rel-L2 {eu:.3e}")
        ratios.append(eg / max(ee, 1e-6))
        # per tensor, CUDA and emulation are two independent noise realisations of the same size (0.1-0.2 in fp32
        # mode from block 0 to the stem, seeded by different roundings; run to run the CUDA value itself moves by
        # that much, tools/diag_2d_det.py): a common cap per tensor, the RATIO is asserted on the median below
        bound = max(f * ee + tol_grad, 4 * floor, 0.3 if precision == "fp32" else 0.9)
        if eg > bound or eu > bound + 0.02:
            failures.append((n, eg, ee, eu, floor))
    med = float(np.median(ratios))
    log(f"[2d step {precision}] {len(ratios)} parameters; median (CUDA error / emulation error) {med:.2f}")
    assert not failures, failures
    assert med < 1.5, med


@pytest.mark.parametrize("nsteps", [1, 3])
def test_graphed_step_2d_matches_eager_step(nsteps):
    """The captured 2-D step (five graphs keyed by index2, draws / beta / lr as device data) against the eager step
    from the same state and the same Python RNG: loss scalars, which parameters own a momentum buffer afterwards
    (= the reached-parameter rule against autograd's own record, note N3), parameters and BatchNorm buffers.
    Not bitwise: two forwards of the SAME input from the SAME state already differ (tools/diag_2d_det.py: the
    statistics' atomics order moves a BatchNorm scale by an fp32 ulp, the tf32 rounding of the stored activation
    turns that into a tf32 ulp in a few elements, and this network amplifies it to 1e-2 of the mask's range by the
    last block), so one step is compared at that noise level and three steps (lr 1e-2, b=4) loosely."""
    from pcrlv2_b200 import train_2d as T2
    from pcrlv2_b200.train_3d import FlatSGD
    crit, cos = torch.nn.MSELoss(), torch.nn.CosineSimilarity()
    batches = [orc.synthetic_batch(4, seed=50 + i, size=(64, 64), local=(32, 32)) for i in range(nsteps)]
    loader = [(b[0], b[1], b[2], b[2], b[3]) for b in batches]
    args = types.SimpleNamespace(lr=1e-2, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    res = {}
    for mode in ("eager", "eager2", "graph"):
        os.environ["PCRL_GRAPH"] = "0" if mode.startswith("eager") else "1"
        try:
            m, sd0 = build2d("fp32")
            opt = FlatSGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
            random.seed(77)
            meters = T2.train_pcrlv2_inner(args, 0, loader, m, opt, crit, cos)
            torch.cuda.synchronize()
            if mode == "graph":
                gs = next(iter(opt._graphed.values()))
                assert isinstance(gs, T2.GraphedStep2d) and len(gs.graphs) >= 1
            name_of = {id(p): n for n, p in m.named_parameters()}
            res[mode] = dict(meters=meters, sd={k: v.detach().cpu().clone() for k, v in m.state_dict().items()},
                             owned={name_of[id(p)] for p in opt.state if "momentum_buffer" in opt.state[p]})
            opt.__dict__.pop("_graphed", None)
        finally:
            os.environ.pop("PCRL_GRAPH", None)
    e, e2, g = res["eager"], res["eager2"], res["graph"]
    log(f"[2d graph, {nsteps} step(s)] meters (cos, mg, local) eager {e['meters']} eager again {e2['meters']} graph {g['meters']}")
    for a, a2, b in zip(e["meters"], e2["meters"], g["meters"]):
        assert abs(a - b) < 3 * abs(a - a2) + (1e-3 if nsteps == 1 else 3e-2), (e["meters"], e2["meters"], g["meters"])
    assert e["owned"] == g["owned"] == e2["owned"], sorted(e["owned"] ^ g["owned"])

    def update_errors(x, y):
        errs = []
        for k, v in x["sd"].items():
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(y["sd"][k]), k
            elif orc.is_param(k) and not orc.is_cancelling(k) and k in x["owned"]:
                du, dv = v.double() - sd0[k].double(), y["sd"][k].double() - sd0[k].double()
                errs.append(((du - dv).norm() / du.norm().clamp_min(1e-30)).item())
        return float(np.median(errs)), max(errs)

    noise, graph = update_errors(e, e2), update_errors(e, g)
    log(f"[2d graph, {nsteps} step(s)] update rel-L2 (median, worst): eager vs eager again {noise[0]:.3e} {noise[1]:.3e}; "
        f"graph vs eager {graph[0]:.3e} {graph[1]:.3e}")
    # the captured step must be no further from the eager step than the eager step is from itself
    assert graph[0] <= 2.0 * noise[0] + 0.02, (graph, noise)


def test_trainer_2d_entry_point(tmp_path):
    """pcrlv2_b200.main --n chest --d 2: epoch loop, LR schedule, the encoder-only checkpoint of train_2d.py:99."""
    from pcrlv2_b200 import main as M
    out = str(tmp_path / "ckpt")
    M.main(["--n", "chest", "--d", "2", "--b", "2", "--epochs", "1", "--synthetic_items", "4", "--workers", "0",
            "--output", out, "--lr", "1e-3"])
    files = os.listdir(out)
    assert files == ["pcrlv2_chest_pretask_0.8_0.pt"], files
    ck = torch.load(os.path.join(out, files[0]), weights_only=False)
    assert ck["epoch"] == 0 and "optimizer" in ck
    keys = list(ck["state_dict"].keys())
    assert keys[0] == "conv1.weight" and "layer4.1.bn2.running_var" in keys and len(keys) == 120   # torchvision resnet18 minus fc
    import torchvision
    ref = torchvision.models.resnet18()
    del ref.fc
    missing = ref.load_state_dict(ck["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
