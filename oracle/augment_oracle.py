"""CPU oracle of the GPU-side input staging (pcrlv2_b200/staging.py, csrc/augment.cu).  TEST INFRASTRUCTURE ONLY.

Restates, in numpy / scipy, the per-item transforms the reference applies through torchio
(/root/reference/data.py:73-89, datasets/lunaDataset.py:28-81).  PARITY UNPINNED for what is torchio's own:
torchio is not installed in this image and the reference pins no version (README.md:7), so its defaults
(parameter ranges, RandomSwap's patch sampling and write order, ZNormalization's Bessel-corrected std) are
restated from torchio's published source (v0.18/0.19 `transforms/augmentation/intensity/*.py`,
`preprocessing/intensity/z_normalization.py`), not checked against an installed copy.  What IS pinned: the
arithmetic of every op against numpy / scipy -- RandomBlur is literally `scipy.ndimage.gaussian_filter`, the
function torchio calls.  RandomAffine (SimpleITK resampling) is not restated and not built.
Every function takes the random PARAMETERS explicitly: the same numbers drive the kernels in the tests."""
import numpy as np
from scipy import ndimage

DEFAULTS = dict(flip_axes=(0,), flip_probability=0.5,           # torchio.RandomFlip()
                blur_std=(0.0, 2.0),                            # torchio.RandomBlur()
                noise_mean=0.0, noise_std=(0.0, 0.25),          # torchio.RandomNoise()
                log_gamma=(-0.3, 0.3),                          # torchio.RandomGamma()
                swap_patch=(8, 4, 4), swap_iterations=100)      # torchio.RandomSwap(patch_size=(8, 4, 4)), data.py:86


def flip(x, mask):
    """x [D,H,W]; bit 0 / 1 / 2 of mask mirrors axis 0 / 1 / 2 (RandomFlip)."""
    for ax in range(3):
        if mask >> ax & 1:
            x = np.flip(x, ax)
    return np.ascontiguousarray(x)


def blur(x, sigma3):
    """RandomBlur: scipy.ndimage.gaussian_filter(x, std) -- float32 in, float32 out, fp64 inside."""
    return ndimage.gaussian_filter(x.astype(np.float32), sigma3)


def noise_gamma(x, noise, noise_std, log_gamma):
    """RandomNoise (mean 0) then RandomGamma: sign(t) |t|^exp(log_gamma) (torchio's rule for negative values)."""
    t = x.astype(np.float32) + np.float32(noise_std) * noise.astype(np.float32)
    return (np.sign(t) * np.abs(t) ** np.float32(np.exp(np.float32(log_gamma)))).astype(np.float32)


def swap(x, corners, patch):
    """RandomSwap: for every iteration take both patches, write the first at the second location, then the
    second at the first location (torchio `_swap` / `_insert` order)."""
    x = x.copy()
    pd, ph, pw = patch
    for c in corners:
        a = x[c[0]:c[0] + pd, c[1]:c[1] + ph, c[2]:c[2] + pw].copy()
        b = x[c[3]:c[3] + pd, c[4]:c[4] + ph, c[5]:c[5] + pw].copy()
        x[c[3]:c[3] + pd, c[4]:c[4] + ph, c[5]:c[5] + pw] = a
        x[c[0]:c[0] + pd, c[1]:c[1] + ph, c[2]:c[2] + pw] = b
    return x


def znorm(x):
    """ZNormalization: (x - mean) / std with torch's default (Bessel-corrected) std, fp64 statistics."""
    x64 = x.astype(np.float64)
    return ((x64 - x64.mean()) / x64.std(ddof=1)).astype(np.float32)


def sample_swap_corners(rng, shape, patch, iterations):
    """torchio RandomSwap location sampling: a uniformly random first patch; second patches are redrawn
    while they lie entirely inside the first one."""
    out = []
    mx = [s - p for s, p in zip(shape, patch)]
    for _ in range(iterations):
        f = [rng.randint(0, m) for m in mx]
        while True:
            s = [rng.randint(0, m) for m in mx]
            inside = all(si >= fi for si, fi in zip(s, f)) and all(si + p <= fi + p for si, fi, p in zip(s, f, patch))
            if not inside:
                break
        out.append(f + s)
    return out
