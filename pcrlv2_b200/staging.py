"""GPU-side input staging for the 3-D pre-training loop.

The reference augments every dataset item on the CPU with torchio (data.py:73-89, datasets/lunaDataset.py:28-81:
RandomFlip, RandomAffine, then RandomBlur, RandomNoise, RandomGamma, RandomSwap((8,4,4)) [global crops only],
ZNormalization -- for 2 global + 6 local volumes per item).  At 300-500 items/s per GPU that CPU pipeline is the
end-to-end bottleneck, so here the RAW crops (the `*_global_k.npy` / `*_local_k.npy` arrays the loader reads,
luna_preprocess.py:141-145) are copied host -> device from pinned memory on a copy stream and augmented by
batched kernels (csrc/augment.cu); the host only draws the random PARAMETERS (same distributions as torchio's
defaults, `PARAMS` below).  RandomAffine (SimpleITK resampling inside torchio) is NOT built:
pass crops that already went through it, or accept its absence.

    aug = GpuAugmenter(device, seed=42)
    for batch in PrefetchLoader(raw_loader, aug):          # (input1, input2, gt1, gt2, [6 local views]) on device
        ...train_pcrlv2_inner consumes exactly this tuple (train_3d.py:109)
"""
import random

import torch

from . import _lib

PARAMS = dict(flip_probability=0.5, blur_std=(0.0, 2.0), noise_std=(0.0, 0.25), log_gamma=(-0.3, 0.3),
              swap_patch=(8, 4, 4), swap_iterations=100)


def _vol(x):
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 5 and x.shape[1] == 1
    b, _, d, h, w = x.shape
    return b, d, h, w


def flip(x, mask):
    b, d, h, w = _vol(x)
    y = torch.empty_like(x)
    _lib.call("pcrl_aug_flip", x, y, mask, b, d, h, w)
    return y


def blur(x, sigma):
    """sigma [B,3] fp32 (axis D, H, W): three separable passes."""
    b, d, h, w = _vol(x)
    t0, t1 = torch.empty_like(x), torch.empty_like(x)
    _lib.call("pcrl_aug_blur_axis", x, t0, sigma, 3, 0, b, d, h, w)
    _lib.call("pcrl_aug_blur_axis", t0, t1, sigma, 3, 1, b, d, h, w)
    _lib.call("pcrl_aug_blur_axis", t1, t0, sigma, 3, 2, b, d, h, w)
    return t0


def noise_gamma(x, noise_std, log_gamma, seed=0, noise=None):
    b, d, h, w = _vol(x)
    y = torch.empty_like(x)
    _lib.call("pcrl_aug_noise_gamma", x, y, noise, noise_std, log_gamma, int(seed) & (2 ** 64 - 1), b, d * h * w)
    return y


def swap_(x, corners, patch):
    """In place.  corners int32 [B, iters, 6]."""
    b, d, h, w = _vol(x)
    assert corners.dtype == torch.int32 and corners.shape[0] == b and corners.shape[2] == 6
    _lib.call("pcrl_aug_swap", x, corners, corners.shape[1], patch[0], patch[1], patch[2], b, d, h, w)
    return x


def znorm(x):
    b, d, h, w = _vol(x)
    y = torch.empty_like(x)
    _lib.call("pcrl_aug_znorm", x, y, b, d * h * w)
    return y


class GpuAugmenter:
    """Draws torchio-default random parameters on the host and applies the transform chain on the device."""

    def __init__(self, device, seed=0, params=None):
        self.dev = torch.device(device)
        self.rng = random.Random(seed)
        self.p = dict(PARAMS, **(params or {}))
        self.step = 0

    def _params(self, b, shape, swap):
        r, p = self.rng, self.p
        mask = [1 if r.random() < p["flip_probability"] else 0 for _ in range(b)]          # RandomFlip(): axis 0
        sigma = [[r.uniform(*p["blur_std"]) for _ in range(3)] for _ in range(b)]
        nstd = [r.uniform(*p["noise_std"]) for _ in range(b)]
        lg = [r.uniform(*p["log_gamma"]) for _ in range(b)]
        corners = None
        if swap:
            pt, mx = p["swap_patch"], [s - q for s, q in zip(shape, p["swap_patch"])]
            corners = []
            for _ in range(b):
                rows = []
                for _ in range(p["swap_iterations"]):
                    f = [r.randint(0, m) for m in mx]
                    while True:
                        s = [r.randint(0, m) for m in mx]
                        if not all(si >= fi and si + q <= fi + q for si, fi, q in zip(s, f, pt)):
                            break
                    rows.append(f + s)
                corners.append(rows)
        return mask, sigma, nstd, lg, corners

    def _to_dev(self, values, dtype):
        return torch.tensor(values, dtype=dtype).pin_memory().to(self.dev, non_blocking=True)

    def spatial(self, x):
        """RandomFlip (the reference's `self.transform`; RandomAffine is not built).  Returns (x', gt) -- gt is
        the spatially transformed copy BEFORE the intensity transforms (lunaDataset.py:38-39)."""
        b = x.shape[0]
        mask = [1 if self.rng.random() < self.p["flip_probability"] else 0 for _ in range(b)]
        return flip(x, self._to_dev(mask, torch.int32))

    def intensity(self, x, swap):
        b, d, h, w = _vol(x)
        _, sigma, nstd, lg, corners = self._params(b, (d, h, w), swap)
        self.step += 1
        y = blur(x, self._to_dev(sigma, torch.float32))
        y = noise_gamma(y, self._to_dev(nstd, torch.float32), self._to_dev(lg, torch.float32),
                        seed=self.rng.getrandbits(63))
        if swap:
            swap_(y, self._to_dev(corners, torch.int32), self.p["swap_patch"])
        return znorm(y)

    def __call__(self, crop1, crop2, local_crops):
        """crop1 / crop2: raw global crops [B,1,D,H,W] on the device, local_crops: list of 6 [B,1,d,h,w].
        Returns the batch tuple of the reference loader: (input1, input2, gt1, gt2, local_inputs)."""
        g1, g2 = self.spatial(crop1), self.spatial(crop2)
        locs = [self.intensity(self.spatial(v), swap=False) for v in local_crops]
        return (self.intensity(g1, True), self.intensity(g2, True), g1, g2, locs)


class PrefetchLoader:
    """Iterates a loader of RAW host batches (crop1, crop2, [local crops]); while the consumer trains on batch i,
    batch i+1 is copied host -> device from pinned memory on a copy stream and augmented there."""

    def __init__(self, loader, augmenter):
        self.loader, self.aug = loader, augmenter
        self.stream = torch.cuda.Stream(device=augmenter.dev)

    def __len__(self):
        return len(self.loader)

    def _stage(self, raw):
        c1, c2, locs = raw
        with torch.cuda.stream(self.stream):
            dev = self.aug.dev

            def up(t):
                t = t.float()
                t = t if t.is_pinned() else t.pin_memory()
                return t.to(dev, non_blocking=True).contiguous()
            out = self.aug(up(c1), up(c2), [up(v) for v in locs])
        return out

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        for raw in it:
            cur = self._hand_over(nxt)
            nxt = self._stage(raw)         # overlaps with the consumer's work on `cur`
            yield cur
        yield self._hand_over(nxt)

    def _hand_over(self, batch):
        """Make the consumer's stream wait for the staging stream and tell the allocator that the tensors
        (allocated on the staging stream) are now in use on the consumer's stream."""
        main = torch.cuda.current_stream(self.aug.dev)
        main.wait_stream(self.stream)
        for t in batch[:4] + tuple(batch[4]):
            t.record_stream(main)
        return batch
