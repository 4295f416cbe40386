"""Op-level parity of every CUDA kernel (through the C ABI) against plain torch fp32 ops on the
same bf16-representable inputs.  Tolerances: fp32-output paths 2e-4 of the reference max
(accumulation order only); bf16-output paths 1.2e-2 (one bf16 rounding of the result)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from pcrlv2_b200 import kernels as K

DEV = "cuda"
TOL32 = 2e-4
TOL16 = 1.2e-2


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def q(t):  # bf16-representable fp32
    return t.to(torch.bfloat16).float()


CONV_SHAPES = [
    (2, 4, 6, 8, 64, 64),
    (1, 8, 8, 4, 128, 128),
    (2, 3, 5, 16, 32, 64),
    (1, 16, 16, 8, 64, 32),
    (3, 2, 2, 2, 256, 128),
    (2, 8, 12, 32, 64, 64),
    (1, 4, 4, 4, 512, 256),
    (1, 6, 10, 16, 128, 64),
    (1, 4, 6, 8, 64, 128),
    (2, 5, 3, 7, 64, 64),
    (1, 4, 16, 32, 64, 64),      # plane of 561 rows, D % 4 == 0, 64 columns: the rolling plane-window kernel
    (2, 8, 20, 24, 32, 64),      # the same kernel with the 32-channel chunk and two window tiles + a ragged third
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_fprop(shape):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(0)
    x = q(torch.randn(n, cin, d, h, w, device=DEV))
    wt = q(torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5)
    ref = F.conv3d(x, wt, padding=1)
    xp = K.pad_ndhwc(x)
    wf, _ = K.pack_conv3_weights(wt)
    y32 = K.conv3d_k3_fprop(xp, wf, out_fp32=True)
    assert rel(K.unpad_ndhwc(y32), ref) < TOL32
    stats = torch.zeros(cout, 2, dtype=torch.float64, device=DEV)
    y16 = K.conv3d_k3_fprop(xp, wf, stats=stats)
    got = K.unpad_ndhwc(y16)
    assert rel(got, ref) < TOL16
    s1 = got.double().sum(dim=(0, 2, 3, 4))
    s2 = (got.double() ** 2).sum(dim=(0, 2, 3, 4))
    assert rel(stats[:, 0], s1) < 1e-4 or (stats[:, 0] - s1).abs().max() < 1e-2
    assert rel(stats[:, 1], s2) < 1e-5
    # per-sample statistics (InstanceNorm)
    st2 = torch.zeros(n, cout, 2, dtype=torch.float64, device=DEV)
    K.conv3d_k3_fprop(xp, wf, stats=st2, per_sample=True)
    assert rel(st2[..., 1], (got.double() ** 2).sum(dim=(2, 3, 4))) < 1e-5


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_dgrad(shape):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(1)
    x = q(torch.randn(n, cin, d, h, w, device=DEV)).requires_grad_(True)
    wt = q(torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5)
    dy = q(torch.randn(n, cout, d, h, w, device=DEV))
    F.conv3d(x, wt, padding=1).backward(dy)
    _, wd = K.pack_conv3_weights(wt)
    dx = K.conv3d_k3_dgrad(K.pad_ndhwc(dy), wd)
    assert rel(K.unpad_ndhwc(dx), x.grad) < TOL16


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_wgrad(shape):
    n, d, h, w, cin, cout = shape
    if cout % 64:
        pytest.skip("weight gradient needs Cout % 64 == 0 (all reference layers satisfy it)")
    torch.manual_seed(2)
    x = q(torch.randn(n, cin, d, h, w, device=DEV))
    wt = q(torch.randn(cout, cin, 3, 3, 3, device=DEV)).requires_grad_(True)
    dy = q(torch.randn(n, cout, d, h, w, device=DEV))
    F.conv3d(x, wt, padding=1).backward(dy)
    gpk = K.conv3d_k3_wgrad(K.pad_ndhwc(dy), K.pad_ndhwc(x))
    assert rel(K.unpack_conv3_wgrad(gpk), wt.grad) < TOL32
    # accumulation into an existing buffer
    gpk2 = K.conv3d_k3_wgrad(K.pad_ndhwc(dy), K.pad_ndhwc(x), out=gpk.clone())
    assert rel(K.unpack_conv3_wgrad(gpk2), 2 * wt.grad) < TOL32


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 16, 16, 16), (3, 4, 6, 2), (2, 64, 64, 32)])
def test_stem_conv(shape):
    n, d, h, w = shape
    torch.manual_seed(3)
    x = torch.randn(n, 1, d, h, w, device=DEV)
    wt = torch.randn(32, 1, 3, 3, 3, device=DEV, requires_grad=True)
    ref = F.conv3d(x, wt, padding=1)
    stats = torch.zeros(32, 2, dtype=torch.float64, device=DEV)
    yp = K.stem_conv_fprop(x, wt.detach(), stats=stats)
    got = K.unpad_ndhwc(yp)
    assert rel(got, ref) < TOL16
    assert yp[:, :, 0].abs().max().item() == 0
    assert rel(stats[:, 1], (got.double() ** 2).sum(dim=(0, 2, 3, 4))) < 1e-5
    dy = q(torch.randn_like(ref))
    ref.backward(dy)
    dw = K.stem_conv_wgrad(K.pad_ndhwc(dy), x)
    assert rel(dw, wt.grad) < TOL32


@pytest.mark.parametrize("shape", [(2, 2, 2, 2, 128, 128), (1, 4, 4, 2, 256, 256), (2, 8, 8, 4, 128, 128),
                                   (1, 2, 3, 5, 512, 512)])
def test_convT(shape):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(4)
    x = q(torch.randn(n, cin, d, h, w, device=DEV)).requires_grad_(True)
    wt = q(torch.randn(cin, cout, 2, 2, 2, device=DEV) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, device=DEV, requires_grad=True)
    ref = F.conv_transpose3d(x, wt, b, stride=2)
    wf, wd = K.pack_convT_weights(wt.detach())
    xp = K.pad_ndhwc(x.detach())
    yp = K.convT_fprop(xp, wf, b.detach())
    assert rel(K.unpad_ndhwc(yp), ref) < TOL16
    assert yp[:, :, 0].abs().max().item() == 0
    g = q(torch.randn_like(ref))
    ref.backward(g)
    dx, dw, db = K.convT_bwd(K.pad_ndhwc(g), xp, wd)
    assert rel(K.unpad_ndhwc(dx), x.grad) < TOL16
    assert dx[:, :, 0].abs().max().item() == 0
    assert rel(K.unpack_convT_wgrad(dw, cin, cout), wt.grad) < TOL32
    assert rel(db, b.grad) < TOL32


def _norm_ref(y, gamma, beta, act, per_sample, slope=None):
    if per_sample:
        z = F.instance_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    else:
        z = F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    if act == "relu":
        return F.relu(z)
    if act == "elu":
        return F.elu(z)
    if act == "prelu":
        return F.prelu(z, slope)
    if act == "sigmoid":
        return torch.sigmoid(z)
    return z


@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("act", ["relu", "elu", "prelu", "sigmoid"])
@pytest.mark.parametrize("mode", ["full", "pool", "avg"])
def test_norm_act(per_sample, act, mode):
    _norm_act_case(per_sample, act, mode, (3, 64, 4, 6, 8))


# rows wider than one 256-thread block (W * C/8 > 256: the 128x128x64 crops of BASELINE configs[3])
# are processed as several segments per row
@pytest.mark.parametrize("per_sample", [False, True])
@pytest.mark.parametrize("mode", ["full", "pool", "avg"])
@pytest.mark.parametrize("dims", [(2, 64, 2, 4, 64), (1, 128, 2, 2, 48), (2, 32, 4, 2, 128)])
def test_norm_act_wide_rows(per_sample, mode, dims):
    _norm_act_case(per_sample, "relu", mode, dims)


def _norm_act_case(per_sample, act, mode, dims):
    n, c, d, h, w = dims
    torch.manual_seed(5)
    y = q(torch.randn(n, c, d, h, w, device=DEV) * 2 + 0.3).requires_grad_(True)
    gamma = (torch.rand(c, device=DEV) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, device=DEV) * 0.2).requires_grad_(True)
    slope = torch.full((c,), 0.25, device=DEV, requires_grad=True) if act == "prelu" else None
    bias = torch.randn(c, device=DEV)
    yp = K.pad_ndhwc(y.detach())
    yp[:, :, 0] = 7.0  # pad rows of a raw conv output are garbage: nothing may depend on them
    G = n if per_sample else 1
    yd = y.detach().double()
    dims = (2, 3, 4) if per_sample else (0, 2, 3, 4)
    stats = torch.stack([yd.sum(dims), (yd ** 2).sum(dims)], -1).reshape(G, c, 2).contiguous()
    count = d * h * w * (1 if per_sample else n)
    rm = torch.zeros(c, device=DEV)
    rv = torch.ones(c, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    scale, shift, mean, invstd = K.norm_finalize(
        stats, count, gamma.detach(), beta.detach(), bias, None if per_sample else rm,
        None if per_sample else rv, None if per_sample else nbt)
    if not per_sample:
        rm_ref, rv_ref = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        F.batch_norm(y.detach() + bias.view(1, -1, 1, 1, 1), rm_ref, rv_ref, None, None, True, 0.1, 1e-5)
        assert rel(rm, rm_ref) < 1e-5 and rel(rv, rv_ref) < 1e-5 and nbt.item() == 1
    a_ref = _norm_ref(y, gamma, beta, act, per_sample, slope)
    a, pool, avg = K.norm_act_fwd(yp, scale, shift, act, slope.detach() if slope is not None else None,
                                  want_full=(mode != "pool"), want_pool=(mode == "pool"),
                                  want_avg=(mode == "avg"), per_sample=per_sample)
    if mode == "pool":
        out_ref = F.max_pool3d(a_ref, 2)
        assert rel(K.unpad_ndhwc(pool), out_ref) < TOL16
        assert pool[:, :, 0].abs().max().item() == 0
    else:
        out_ref = a_ref
        assert rel(K.unpad_ndhwc(a), a_ref) < TOL16
        assert a[:, :, 0].abs().max().item() == 0
    loss_terms = []
    g1 = q(torch.randn_like(out_ref))
    loss_terms.append((out_ref * g1).sum())
    g2 = gavg = None
    if mode == "avg":
        assert rel(avg / (d * h * w), a_ref.mean(dim=(2, 3, 4))) < 2e-3
        g2 = q(torch.randn_like(a_ref))
        gavg = torch.randn(n, c, device=DEV)
        loss_terms.append((a_ref * g2).sum())
        loss_terms.append((a_ref.mean(dim=(2, 3, 4)) * gavg).sum())
    sum(loss_terms).backward()
    dy, sums = K.norm_act_bwd(yp, K.pad_ndhwc(g1), K.pad_ndhwc(g2) if g2 is not None else None, gavg,
                              scale, shift, mean, invstd, gamma.detach(), act,
                              slope.detach() if slope is not None else None, pool=(mode == "pool"),
                              per_sample=per_sample)
    assert rel(K.unpad_ndhwc(dy), y.grad) < 2e-2
    assert dy[:, :, 0].abs().max().item() == 0
    assert rel(sums[..., 0].sum(0), beta.grad) < 2e-3
    assert rel(sums[..., 1].sum(0), gamma.grad) < 2e-3
    if act == "prelu":
        assert rel(sums[..., 2].sum(0), slope.grad) < 2e-3


@pytest.mark.parametrize("shape", [(2, 4, 6, 8, 64), (1, 8, 8, 4, 128), (3, 2, 2, 2, 256), (2, 16, 16, 16, 64)])
@pytest.mark.parametrize("with_final", [False, True])
def test_heads(shape, with_final):
    """Conv3d(C->1,k3,p1) [+ Conv3d(C->1,k1)] factored as tensor-core GEMM + 27-point gather."""
    n, d, h, w, c = shape
    torch.manual_seed(6)
    a = q(torch.randn(n, c, d, h, w, device=DEV)).requires_grad_(True)
    w3 = q(torch.randn(1, c, 3, 3, 3, device=DEV) / (27 * c) ** 0.5).requires_grad_(True)
    b3 = torch.randn(1, device=DEV, requires_grad=True)
    w1 = q(torch.randn(1, c, 1, 1, 1, device=DEV) / c ** 0.5).requires_grad_(True)
    b1 = torch.randn(1, device=DEV, requires_grad=True)
    y1_ref = F.conv3d(a, w3, b3, padding=1)
    y0_ref = F.conv3d(a, w1, b1)
    ap = K.pad_ndhwc(a.detach())
    wext, wext_t = K.head_pack_weights(w3.detach(), w1.detach() if with_final else None)
    stats = torch.zeros(1, 1, 2, dtype=torch.float64, device=DEV)
    y1, y0 = K.head_fwd(ap, wext, b3.detach(), b1.detach() if with_final else None, stats)
    assert rel(y1, y1_ref) < TOL32
    assert rel(stats[0, 0, 0:1], y1_ref.double().sum().reshape(1)) < 1e-4 or abs(stats[0, 0, 0].item() - y1_ref.double().sum().item()) < 1e-2
    assert rel(stats[0, 0, 1:2], (y1_ref.double() ** 2).sum().reshape(1)) < 1e-4
    st_n = torch.zeros(n, 1, 2, dtype=torch.float64, device=DEV)
    K.head_fwd(ap, wext, b3.detach(), None, st_n, per_sample=True)
    assert rel(st_n[:, 0, 1], (y1_ref.double() ** 2).sum(dim=(1, 2, 3, 4))) < 1e-4
    dy1 = q(torch.randn_like(y1_ref))
    dy0 = q(torch.randn_like(y0_ref))
    loss = (y1_ref * dy1).sum()
    if with_final:
        assert rel(y0, y0_ref) < TOL32
        loss = loss + (y0_ref * dy0).sum()
    loss.backward()
    da, dwext = K.head_bwd(ap, dy1, dy0 if with_final else None, wext_t)
    assert rel(K.unpad_ndhwc(da), a.grad) < TOL16
    assert da[:, :, 0].abs().max().item() == 0
    assert rel(dwext[:, :27].reshape(1, c, 3, 3, 3), w3.grad) < TOL32
    if with_final:
        assert rel(dwext[:, 27].reshape(1, c, 1, 1, 1), w1.grad) < TOL32


@pytest.mark.parametrize("per_sample", [False, True])
def test_chan1_norm_sigmoid(per_sample):
    n, d, h, w = 3, 4, 6, 8
    torch.manual_seed(9)
    y = (torch.randn(n, 1, d, h, w, device=DEV) * 1.7 + 0.4).requires_grad_(True)
    gamma = torch.tensor([1.3], device=DEV, requires_grad=True)
    beta = torch.tensor([-0.2], device=DEV, requires_grad=True)
    if per_sample:
        ref = torch.sigmoid(F.instance_norm(y, None, None, gamma, beta, True, 0.1, 1e-5))
    else:
        ref = torch.sigmoid(F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5))
    yd = y.detach().double()
    dims = (1, 2, 3, 4) if per_sample else (0, 1, 2, 3, 4)
    G = n if per_sample else 1
    stats = torch.stack([yd.sum(dims), (yd ** 2).sum(dims)], -1).reshape(G, 1, 2).contiguous()
    count = d * h * w * (1 if per_sample else n)
    scale, shift, mean, invstd = K.norm_finalize(stats, count, gamma.detach(), beta.detach())
    mask = K.chan1_sigmoid_fwd(y.detach().contiguous(), scale, shift, per_sample)
    assert rel(mask, ref) < 1e-5
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    dy, sums = K.chan1_sigmoid_bwd(y.detach().contiguous(), mask, g, mean, invstd, gamma.detach(), per_sample)
    assert rel(dy, y.grad) < 1e-4
    assert rel(sums[:, 1].sum().reshape(1), gamma.grad) < 1e-4
    assert rel(sums[:, 0].sum().reshape(1), beta.grad) < 1e-4


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 16, 16, 16), (3, 4, 6, 2), (2, 64, 64, 32)])
def test_stem_wgrad_gemm(shape):
    n, d, h, w = shape
    torch.manual_seed(3)
    x = q(torch.randn(n, 1, d, h, w, device=DEV))
    wt = torch.randn(32, 1, 3, 3, 3, device=DEV, requires_grad=True)
    ref = F.conv3d(x, wt, padding=1)
    dy = q(torch.randn_like(ref))
    ref.backward(dy)
    dw = K.stem_conv_wgrad_gemm(K.pad_ndhwc(dy), x)
    assert rel(dw, wt.grad) < TOL32


@pytest.mark.parametrize("dims", [(300, 128, 256), (128, 64, 64), (1000, 512, 128), (77, 256, 512)])
def test_gemms(dims):
    rows, k, cols = dims
    torch.manual_seed(7)
    a = q(torch.randn(rows, k, device=DEV))
    b = q(torch.randn(cols, k, device=DEV))
    bias = torch.randn(cols, device=DEV)
    c = K.gemm_nt(a.to(torch.bfloat16), b.to(torch.bfloat16), bias)
    assert rel(c, a @ b.t() + bias) < TOL32
    b2 = q(torch.randn(rows, cols, device=DEV))
    c2 = K.gemm_tn(a.to(torch.bfloat16), b2.to(torch.bfloat16))
    assert rel(c2, a.t() @ b2) < TOL32
    ct = K.gemm_nt(a.to(torch.bfloat16), b.to(torch.bfloat16), None, out_fp32="transposed")
    assert rel(ct, (a @ b.t()).t()) < TOL32
    # 32-wide operands (head GEMMs): K = 32 and Q = 32
    a32 = q(torch.randn(rows, 32, device=DEV))
    c3 = K.gemm_nt(a32.to(torch.bfloat16), q(torch.randn(cols, 32, device=DEV)).to(torch.bfloat16))
    assert c3.shape == (rows, cols)
    b32 = q(torch.randn(cols, 32, device=DEV))
    assert rel(K.gemm_nt(a32.to(torch.bfloat16), b32.to(torch.bfloat16)), a32 @ b32.t()) < TOL32
    assert rel(K.gemm_tn(a.to(torch.bfloat16), a32.to(torch.bfloat16)), a.t() @ a32) < TOL32


def test_sgd_flat():
    torch.manual_seed(8)
    sizes = [5, 1000, 64, 27 * 64 * 64, 3]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    tot = off[-1]
    p = torch.randn(tot, device=DEV)
    g = torch.randn(tot, device=DEV)
    buf = torch.randn(tot, device=DEV)
    active = torch.tensor([1, 0, 1, 1, 1], dtype=torch.int32, device=DEV)
    first = torch.tensor([0, 0, 1, 0, 0], dtype=torch.int32, device=DEV)
    seg = torch.tensor(off, dtype=torch.long, device=DEV)
    p0, b0 = p.clone(), buf.clone()
    K.sgd_flat(p, g, buf, seg, active, first, 0.01, 0.9, 1e-4)
    for i, s in enumerate(sizes):
        sl = slice(off[i], off[i + 1])
        if not active[i]:
            assert torch.equal(p[sl], p0[sl]) and torch.equal(buf[sl], b0[sl])
            continue
        d_ = g[sl] + 1e-4 * p0[sl]
        m = d_ if first[i] else 0.9 * b0[sl] + d_
        assert rel(buf[sl], m) < 1e-6 and rel(p[sl], p0[sl] - 0.01 * m) < 1e-6


@pytest.mark.parametrize("shape", [(2, 4, 4, 4, 64, 128), (1, 8, 8, 8, 128, 64), (2, 2, 6, 4, 128, 256)])
def test_dgrad_unshuffled_feeds_convT_bwd(shape):
    """conv data gradient written coarse-major == unshuffle(conv data gradient), and the
    ConvTranspose gradient GEMMs on it match autograd."""
    n, d, h, w, cin, cout = shape          # fine dims of the conv; cin = channels of the convT output
    torch.manual_seed(11)
    xc = q(torch.randn(n, cin, d // 2, h // 2, w // 2, device=DEV)).requires_grad_(True)
    wt = q(torch.randn(cin, cin, 2, 2, 2, device=DEV) / cin ** 0.5).requires_grad_(True)
    bt = torch.randn(cin, device=DEV, requires_grad=True)
    wc = q(torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5)
    up = F.conv_transpose3d(xc, wt, bt, stride=2)
    y = F.conv3d(up, wc, padding=1)
    dy = q(torch.randn_like(y))
    y.backward(dy)
    _, wd = K.pack_conv3_weights(wc)
    _, wtd = K.pack_convT_weights(wt.detach())
    scratch, colsum = K.conv3d_k3_dgrad_unshuffled(K.pad_ndhwc(dy), wd)
    dxc, dwt = K.convT_bwd_from_scratch(scratch, K.pad_ndhwc(xc.detach()), wtd)
    assert rel(K.unpad_ndhwc(dxc), xc.grad) < 2e-2
    assert rel(K.unpack_convT_wgrad(dwt, cin, cin), wt.grad) < 2e-2
    assert rel(colsum[:, 0], bt.grad) < 2e-2


# ------------------------------------------------------------------------------------------------
# fp32 storage / TF32 tensor-core operands (precision="fp32").  Inputs are made TF32-representable
# (10-bit mantissa) so that the tensor-core products are exact and only accumulation order differs.
def q32(t):
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


F32 = torch.float32


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3d_fp32_tf32(shape):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(20)
    x = q32(torch.randn(n, cin, d, h, w, device=DEV)).requires_grad_(True)
    wt = q32(torch.randn(cout, cin, 3, 3, 3, device=DEV) / (27 * cin) ** 0.5).requires_grad_(True)
    ref = F.conv3d(x, wt, padding=1)
    dy = q32(torch.randn_like(ref))
    ref.backward(dy)
    xp = K.pad_ndhwc(x.detach(), F32)
    wf, wd = K.pack_conv3_weights(wt.detach(), dtype=F32)
    stats = torch.zeros(cout, 2, dtype=torch.float64, device=DEV)
    y = K.conv3d_k3_fprop(xp, wf, stats=stats)
    assert y.dtype == F32
    got = K.unpad_ndhwc(y)
    assert rel(got, ref) < TOL32
    assert rel(stats[:, 1], (got.double() ** 2).sum(dim=(0, 2, 3, 4))) < 1e-5
    dyp = K.pad_ndhwc(dy, F32)
    assert rel(K.unpad_ndhwc(K.conv3d_k3_dgrad(dyp, wd)), x.grad) < TOL32
    if cout % 64 == 0:
        assert rel(K.unpack_conv3_wgrad(K.conv3d_k3_wgrad(dyp, xp)), wt.grad) < TOL32


@pytest.mark.parametrize("shape", [(2, 2, 2, 2, 128, 128), (2, 8, 8, 4, 128, 128), (1, 2, 3, 5, 512, 512)])
def test_convT_fp32(shape):
    n, d, h, w, cin, cout = shape
    torch.manual_seed(21)
    x = q32(torch.randn(n, cin, d, h, w, device=DEV)).requires_grad_(True)
    wt = q32(torch.randn(cin, cout, 2, 2, 2, device=DEV) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, device=DEV, requires_grad=True)
    ref = F.conv_transpose3d(x, wt, b, stride=2)
    wf, wd = K.pack_convT_weights(wt.detach(), dtype=F32)
    xp = K.pad_ndhwc(x.detach(), F32)
    yp = K.convT_fprop(xp, wf, b.detach())
    # outputs that feed the next tensor-core kernel are stored tf32-rounded (2^-11 relative)
    assert rel(K.unpad_ndhwc(yp), ref) < 6e-4
    assert yp[:, :, 0].abs().max().item() == 0
    g = q32(torch.randn_like(ref))
    ref.backward(g)
    dx, dw, db = K.convT_bwd(K.pad_ndhwc(g, F32), xp, wd)
    assert rel(K.unpad_ndhwc(dx), x.grad) < 6e-4
    assert rel(K.unpack_convT_wgrad(dw, cin, cout), wt.grad) < TOL32
    assert rel(db, b.grad) < TOL32
    # fused path: conv data gradient written coarse-major
    wc = q32(torch.randn(64, cout, 3, 3, 3, device=DEV) / (27 * cout) ** 0.5)
    x.grad = None; wt.grad = None; b.grad = None
    up = F.conv_transpose3d(x, wt, b, stride=2)
    y = F.conv3d(up, wc, padding=1)
    dy = q32(torch.randn_like(y))
    y.backward(dy)
    _, wcd = K.pack_conv3_weights(wc, dtype=F32)
    scratch, colsum = K.conv3d_k3_dgrad_unshuffled(K.pad_ndhwc(dy, F32), wcd)
    dxc, dwt = K.convT_bwd_from_scratch(scratch, xp, wd)
    # (the intermediate gradient is not TF32-representable: one TF32 rounding of an operand)
    assert rel(K.unpad_ndhwc(dxc), x.grad) < 2e-3
    assert rel(K.unpack_convT_wgrad(dwt, cin, cout), wt.grad) < 2e-3
    # column sums are taken over the stored (tf32-rounded) gradient: ~2^-12 / sqrt(terms) of its rms
    assert rel(colsum[:, 0], b.grad) < 6e-4


@pytest.mark.parametrize("mode", ["full", "pool", "avg"])
def test_norm_act_fp32(mode):
    n, c, d, h, w = 3, 64, 4, 6, 8
    torch.manual_seed(22)
    y = (torch.randn(n, c, d, h, w, device=DEV) * 2 + 0.3).requires_grad_(True)
    gamma = (torch.rand(c, device=DEV) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, device=DEV) * 0.2).requires_grad_(True)
    yp = K.pad_ndhwc(y.detach(), F32)
    yd = y.detach().double()
    stats = torch.stack([yd.sum((0, 2, 3, 4)), (yd ** 2).sum((0, 2, 3, 4))], -1).reshape(1, c, 2).contiguous()
    scale, shift, mean, invstd = K.norm_finalize(stats, n * d * h * w, gamma.detach(), beta.detach())
    a_ref = F.relu(F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5))
    a, pool, avg = K.norm_act_fwd(yp, scale, shift, "relu", None, want_full=(mode != "pool"),
                                  want_pool=(mode == "pool"), want_avg=(mode == "avg"))
    out_ref = F.max_pool3d(a_ref, 2) if mode == "pool" else a_ref
    got = pool if mode == "pool" else a
    assert got.dtype == F32 and rel(K.unpad_ndhwc(got), out_ref) < 6e-4   # stored tf32-rounded
    assert got[:, :, 0].abs().max().item() == 0
    g1 = torch.randn_like(out_ref)
    loss = (out_ref * g1).sum()
    gavg = None
    if mode == "avg":
        # the pooled sum is over the stored (tf32-rounded) activations
        assert rel(avg / (d * h * w), a_ref.mean(dim=(2, 3, 4))) < 2e-4
        gavg = torch.randn(n, c, device=DEV)
        loss = loss + (a_ref.mean(dim=(2, 3, 4)) * gavg).sum()
    loss.backward()
    dy, sums = K.norm_act_bwd(yp, K.pad_ndhwc(g1, F32), None, gavg, scale, shift, mean, invstd,
                              gamma.detach(), "relu", None, pool=(mode == "pool"))
    assert rel(K.unpad_ndhwc(dy), y.grad) < 6e-4
    assert rel(sums[..., 1].sum(0), gamma.grad) < 1e-4


def test_heads_and_stem_fp32():
    n, d, h, w, c = 2, 4, 6, 8, 64
    torch.manual_seed(23)
    a = q32(torch.randn(n, c, d, h, w, device=DEV)).requires_grad_(True)
    w3 = q32(torch.randn(1, c, 3, 3, 3, device=DEV) / (27 * c) ** 0.5).requires_grad_(True)
    b3 = torch.randn(1, device=DEV)
    w1 = q32(torch.randn(1, c, 1, 1, 1, device=DEV) / c ** 0.5).requires_grad_(True)
    b1 = torch.randn(1, device=DEV)
    y1_ref, y0_ref = F.conv3d(a, w3, b3, padding=1), F.conv3d(a, w1, b1)
    ap = K.pad_ndhwc(a.detach(), F32)
    wext, wext_t = K.head_pack_weights(w3.detach(), w1.detach(), dtype=F32)
    y1, y0 = K.head_fwd(ap, wext, b3, b1)
    assert rel(y1, y1_ref) < TOL32 and rel(y0, y0_ref) < TOL32
    dy1, dy0 = q32(torch.randn_like(y1_ref)), q32(torch.randn_like(y0_ref))
    ((y1_ref * dy1).sum() + (y0_ref * dy0).sum()).backward()
    da, dwext = K.head_bwd(ap, dy1, dy0, wext_t)
    assert rel(K.unpad_ndhwc(da), a.grad) < 6e-4
    assert rel(dwext[:, :27].reshape(1, c, 3, 3, 3), w3.grad) < TOL32
    assert rel(dwext[:, 27].reshape(1, c, 1, 1, 1), w1.grad) < TOL32
    # stem
    x = q32(torch.randn(n, 1, 8, 8, 8, device=DEV))
    wt = torch.randn(32, 1, 3, 3, 3, device=DEV, requires_grad=True)
    ref = F.conv3d(x, wt, padding=1)
    yp = K.stem_conv_fprop(x, wt.detach(), dtype=F32)
    assert yp.dtype == F32 and rel(K.unpad_ndhwc(yp), ref) < 6e-4
    dy = q32(torch.randn_like(ref))
    ref.backward(dy)
    assert rel(K.stem_conv_wgrad_gemm(K.pad_ndhwc(dy, F32), x), wt.grad) < TOL32


# ------------------------------------------------------------------------------ heads / losses (csrc/losses.cu)
@pytest.mark.parametrize("shape", [(32, 256), (192, 64), (2, 128), (7, 40)])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_bn1d(shape, relu, training):
    from pcrlv2_b200 import functional as Fn
    b, c = shape
    torch.manual_seed(30)
    bn = torch.nn.BatchNorm1d(c).to(DEV)
    ref = torch.nn.BatchNorm1d(c).to(DEV)
    with torch.no_grad():
        for m in (bn, ref):
            m.weight.copy_(torch.linspace(0.5, 1.5, c))
            m.bias.copy_(torch.linspace(-0.3, 0.3, c))
            m.running_mean.copy_(torch.linspace(-0.1, 0.1, c))
            m.running_var.copy_(torch.linspace(0.8, 1.2, c))
    bn.train(training)
    ref.train(training)
    x = (torch.randn(b, c, device=DEV) * 1.5 + 0.2).requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    y = Fn.batch_norm1d(x, bn, relu=relu)
    y_ref = ref(x2)
    if relu:
        y_ref = torch.relu(y_ref)
    assert rel(y, y_ref) < 1e-5
    assert rel(bn.running_mean, ref.running_mean) < 1e-5 and rel(bn.running_var, ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)
    g = torch.randn_like(y_ref)
    y.backward(g)
    y_ref.backward(g)
    assert rel(x.grad, x2.grad) < 2e-4
    assert rel(bn.weight.grad, ref.weight.grad) < 1e-4 and rel(bn.bias.grad, ref.bias.grad) < 1e-4


@pytest.mark.parametrize("shape", [(32, 256, 512), (192, 128, 64), (2, 64, 128), (5, 33, 70)])
def test_linear(shape):
    from pcrlv2_b200 import functional as Fn
    b, k, j = shape
    torch.manual_seed(31)
    lin = torch.nn.Linear(k, j).to(DEV)
    x = torch.randn(b, k, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = Fn.linear(x, lin)
        y_ref = F.linear(x2, lin.weight, lin.bias)
        assert rel(y, y_ref) < 1e-5
        g = torch.randn_like(y_ref)
        y.backward(g)
        gw, gb = lin.weight.grad.clone(), lin.bias.grad.clone()
        lin.zero_grad()
        y_ref.backward(g)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert rel(x.grad, x2.grad) < 1e-5
    assert rel(gw, lin.weight.grad) < 1e-5 and rel(gb, lin.bias.grad) < 1e-5


@pytest.mark.parametrize("shape", [(32, 256), (192, 64), (2, 128), (3, 50)])
def test_cosine_mean(shape):
    from pcrlv2_b200 import functional as Fn
    b, c = shape
    torch.manual_seed(32)
    x = torch.randn(b, c, device=DEV, requires_grad=True)
    yv = torch.randn(b, c, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    out = Fn.cosine_mean(x, yv, 1e-8, -0.5)
    ref = -0.5 * torch.nn.CosineSimilarity()(x2, yv.detach()).mean()
    assert abs(out.item() - ref.item()) < 1e-6
    (out * 3.0).backward()
    (ref * 3.0).backward()
    assert rel(x.grad, x2.grad) < 1e-5
    assert yv.grad is None


@pytest.mark.parametrize("n", [(2, 1, 64, 64, 32), (3, 1, 5, 7, 3), (1, 1, 16, 16, 16)])
def test_mse_and_sigmoid(n):
    from pcrlv2_b200 import functional as Fn
    torch.manual_seed(33)
    z = torch.randn(n, device=DEV, requires_grad=True)
    z2 = z.detach().clone().requires_grad_(True)
    t = torch.rand(n, device=DEV)
    p = Fn.sigmoid(z)
    loss = Fn.mse_loss(p, t)
    p_ref = torch.sigmoid(z2)
    loss_ref = torch.nn.MSELoss()(p_ref, t)
    assert rel(p, p_ref) < 1e-6
    assert abs(loss.item() - loss_ref.item()) < 1e-6 * max(1.0, abs(loss_ref.item()))
    (loss * 0.7).backward()
    (loss_ref * 0.7).backward()
    assert rel(z.grad, z2.grad) < 1e-5


@pytest.mark.parametrize("shape,sf", [((2, 1, 16, 16, 8), 4), ((2, 1, 32, 32, 16), 2), ((1, 1, 3, 5, 2), 2), ((3, 1, 4, 4, 4), 4)])
def test_upsample_trilinear(shape, sf):
    from pcrlv2_b200 import functional as Fn
    torch.manual_seed(34)
    x = torch.randn(shape, device=DEV, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    y = Fn.upsample_trilinear(x, sf)
    y_ref = F.interpolate(x2, scale_factor=sf, mode="trilinear")
    assert y.shape == y_ref.shape
    assert rel(y, y_ref) < 1e-6
    g = torch.randn_like(y_ref)
    y.backward(g)
    y_ref.backward(g)
    assert rel(x.grad, x2.grad) < 1e-5
    # closed form of SURVEY 8c: trilinear x2 of [0,1,2,3] along one axis
    ramp = torch.arange(4, dtype=torch.float32, device=DEV).view(1, 1, 1, 1, 4).expand(1, 1, 2, 2, 4).contiguous()
    up = Fn.upsample_trilinear(ramp, 2)[0, 0, 0, 0]
    assert torch.allclose(up, torch.tensor([0, .25, .75, 1.25, 1.75, 2.25, 2.75, 3], device=DEV))
