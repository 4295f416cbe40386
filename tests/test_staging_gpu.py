"""GPU-side input staging (pcrlv2_b200/staging.py, csrc/augment.cu) against the numpy / scipy
restatement of the torchio transforms (oracle/augment_oracle.py) with the SAME explicit random parameters.
Flip and patch swap are index work: bit-exact.  Blur (fp64 weights / accumulation on both sides), noise+gamma
and z-normalisation: 2e-6 of the value range."""
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import augment_oracle as ao  # noqa: E402

if torch.cuda.is_available():
    from pcrlv2_b200 import staging as S

SHAPES = [(3, 64, 64, 32), (5, 16, 16, 16), (2, 24, 40, 8)]


def vols(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((shape[0], 1) + tuple(shape[1:]), generator=g)       # HU-normalised crops live in [0, 1]


def dev_i(v):
    return torch.tensor(v, dtype=torch.int32, device="cuda")


def dev_f(v):
    return torch.tensor(v, dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("shape", SHAPES)
def test_flip_and_swap_are_bit_exact(shape):
    x = vols(shape, 1)
    b = shape[0]
    masks = [i % 8 for i in range(b)]
    y = S.flip(x.cuda(), dev_i(masks)).cpu()
    for i in range(b):
        assert np.array_equal(y[i, 0].numpy(), ao.flip(x[i, 0].numpy(), masks[i]))
    rng = random.Random(3)
    patch = (8, 4, 4)
    corners = [ao.sample_swap_corners(rng, shape[1:], patch, 100) for _ in range(b)]
    z = S.swap_(x.cuda().clone(), dev_i(corners), patch).cpu()
    for i in range(b):
        assert np.array_equal(z[i, 0].numpy(), ao.swap(x[i, 0].numpy(), corners[i], patch))
    # overlapping patches: the write order of torchio (second patch last) decides the overlap
    ov = [[[0, 0, 0, 4, 2, 2]] for _ in range(b)]
    z = S.swap_(x.cuda().clone(), dev_i(ov), patch).cpu()
    for i in range(b):
        assert np.array_equal(z[i, 0].numpy(), ao.swap(x[i, 0].numpy(), ov[i], patch))


@pytest.mark.parametrize("shape", SHAPES)
def test_blur_matches_scipy_gaussian_filter(shape):
    x = vols(shape, 2)
    b = shape[0]
    rng = random.Random(5)
    sig = [[rng.uniform(0, 2) for _ in range(3)] for _ in range(b)]
    sig[0] = [0.0, 1.3, 0.0]                      # zero sigma = identity along that axis
    y = S.blur(x.cuda(), dev_f(sig)).cpu()
    for i in range(b):
        ref = ao.blur(x[i, 0].numpy(), sig[i])
        assert np.abs(y[i, 0].numpy() - ref).max() < 2e-6, (i, sig[i])


@pytest.mark.parametrize("shape", SHAPES)
def test_noise_gamma_and_znorm(shape):
    x = vols(shape, 3)
    b = shape[0]
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(x.shape, generator=g)
    nstd = [0.25 * (i + 1) / b for i in range(b)]
    lg = [-0.3 + 0.6 * i / max(b - 1, 1) for i in range(b)]
    y = S.noise_gamma(x.cuda(), dev_f(nstd), dev_f(lg), noise=noise.cuda()).cpu()
    for i in range(b):
        ref = ao.noise_gamma(x[i, 0].numpy(), noise[i, 0].numpy(), nstd[i], lg[i])
        assert np.abs(y[i, 0].numpy() - ref).max() < 4e-6 * max(1.0, np.abs(ref).max()), i
    z = S.znorm(y.cuda()).cpu()
    for i in range(b):
        ref = ao.znorm(y[i, 0].numpy())
        assert np.abs(z[i, 0].numpy() - ref).max() < 1e-5
        assert abs(z[i].mean().item()) < 1e-5 and abs(z[i].std().item() - 1.0) < 1e-5
    # device-generated noise: deterministic per seed, standard normal statistics
    a = S.noise_gamma(x.cuda(), dev_f([1.0] * b), dev_f([0.0] * b), seed=7)
    a2 = S.noise_gamma(x.cuda(), dev_f([1.0] * b), dev_f([0.0] * b), seed=7)
    a3 = S.noise_gamma(x.cuda(), dev_f([1.0] * b), dev_f([0.0] * b), seed=8)
    assert torch.equal(a, a2) and not torch.equal(a, a3)
    n = (a - x.cuda()).flatten().double()
    assert abs(n.mean().item()) < 0.02 and abs(n.std().item() - 1.0) < 0.02
    assert abs((n ** 3).mean().item()) < 0.05 and abs((n ** 4).mean().item() - 3.0) < 0.1


def test_prefetch_loader_feeds_the_trainer():
    """Raw crops (host) -> PrefetchLoader + GpuAugmenter -> train_pcrlv2_inner: the batch contract of the
    reference loader (lunaDataset.py:79-81) is met and the step runs."""
    from oracle import pcrlv2_oracle as orc
    from pcrlv2_b200 import train_3d as T
    from pcrlv2_b200.models import PCRLv23d
    bsz = 4
    raw = []
    for i in range(3):
        g = torch.Generator().manual_seed(20 + i)
        raw.append((torch.rand(bsz, 1, 64, 64, 32, generator=g), torch.rand(bsz, 1, 64, 64, 32, generator=g),
                    [torch.rand(bsz, 1, 16, 16, 16, generator=g) for _ in range(6)]))
    aug = S.GpuAugmenter("cuda", seed=42)
    batches = list(S.PrefetchLoader(raw, aug))
    assert len(batches) == 3
    x1, x2, gt1, gt2, loc = batches[0]
    assert tuple(x1.shape) == (bsz, 1, 64, 64, 32) == tuple(gt1.shape) and len(loc) == 6
    assert tuple(loc[0].shape) == (bsz, 1, 16, 16, 16)
    assert abs(x1[0].mean().item()) < 1e-4 and abs(x1[0].std().item() - 1) < 1e-4        # ends in ZNormalization
    assert 0 <= gt1.min().item() and gt1.max().item() <= 1                               # gt: flipped raw crop
    # gt is the raw crop up to the flip
    r0 = raw[0][0][0, 0]
    assert torch.equal(gt1[0, 0].cpu(), r0) or torch.equal(gt1[0, 0].cpu(), torch.flip(r0, (0,)))
    # same seed -> same parameters -> same batch
    again = list(S.PrefetchLoader(raw, S.GpuAugmenter("cuda", seed=42)))
    assert torch.equal(again[1][0], batches[1][0]) and torch.equal(again[2][4][3], batches[2][4][3])
    m = PCRLv23d()
    m.load_state_dict(orc.init_state(0))
    m = m.cuda().train()
    opt = T.FlatSGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    args = types.SimpleNamespace(lr=1e-3, momentum=0.9, weight_decay=1e-4, amp=False, epochs=240)
    random.seed(1)
    mg, local = T.train_pcrlv2_inner(args, 0, S.PrefetchLoader(raw, S.GpuAugmenter("cuda", seed=1)), m, opt,
                                     torch.nn.MSELoss(), torch.nn.CosineSimilarity())
    assert np.isfinite(mg) and np.isfinite(local) and 0 < mg < 1


def test_crop_generator_pieces_bit_exact():
    """pcrlv2_b200/preprocess.py against the reference's own loops (oracle/preprocess_oracle.py restates
    luna_preprocess.py:132-135, 213-241, 295-320): HU window, depth scan (t_img, d_img, lung sum), IoU."""
    from oracle import preprocess_oracle as po
    from pcrlv2_b200 import preprocess as P
    g = np.random.default_rng(0)
    hu = g.integers(-1400, 1600, size=(40, 36, 20)).astype(np.float32)
    w = P.hu_window(torch.from_numpy(hu).cuda()).cpu().numpy()
    assert np.array_equal(w, po.hu_window(hu).astype(np.float32))
    # small case against the literal nested loops, large case against their vectorised form
    crop = g.random((12, 10, 8 + 3)).astype(np.float32)
    t, d, s = P.depth_scan(torch.from_numpy(crop).cuda(), 8)
    tl, dl = po.depth_scan_loops(crop, 8)
    assert np.array_equal(t.cpu().numpy(), tl.astype(np.float32)) and np.array_equal(d.cpu().numpy(), dl)
    tv, dv = po.depth_scan(crop, 8)
    assert np.array_equal(tv.astype(np.float32), tl.astype(np.float32)) and np.array_equal(dv, dl)
    crop = g.random((64, 64, 32 + 3)).astype(np.float32)
    t, d, s = P.depth_scan(torch.from_numpy(crop).cuda(), 32)
    tv, dv = po.depth_scan(crop, 32)
    assert np.array_equal(t.cpu().numpy(), tv.astype(np.float32)) and np.array_equal(d.cpu().numpy(), dv)
    assert abs(s.item() - float(dv.astype(np.float64).sum())) < 1e-6
    assert P.accept_crop(torch.from_numpy(crop).cuda(), 64, 64, 32) == bool(np.sum(dv) <= 0.15 * 64 * 64 * 32)
    for b1, b2 in [((0, 64, 0, 64, 0, 32), (16, 80, 8, 72, 4, 36)), ((0, 96, 0, 96, 0, 64), (100, 164, 0, 64, 0, 32))]:
        assert P.cal_iou(b1, b2) == po.cal_iou(b1, b2)
