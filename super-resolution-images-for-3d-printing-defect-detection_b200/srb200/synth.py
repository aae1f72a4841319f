"""Seeded synthetic image batches.

HR images are smooth-plus-texture float32 RGB in [0, 1] so that PSNR/SSIM are not
degenerate; LR images are an exact box-filter (area) down-sample, which is what
``cv2.resize(..., INTER_AREA)`` computes for integer factors.  Seeds start at the
reference's ``RANDOM_SEED = 42`` (/root/reference/SRModels/constants.py:14).
"""
from __future__ import annotations

import numpy as np

RANDOM_SEED = 42


def _gauss_taps(sigma):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    g = np.exp(-0.5 * (x / sigma) ** 2)
    return g / g.sum()


def _blur_axis(a, taps, axis):
    r = len(taps) // 2
    pad = [(0, 0)] * a.ndim
    pad[axis] = (r, r)
    ap = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    for k, t in enumerate(taps):
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(k, k + a.shape[axis])
        out += t * ap[tuple(sl)]
    return out


def hr_image(h, w, index=0, sigma=3.0, texture=0.05, channels=3):
    """One HR image: clip(rescale(lowpass(noise)) + texture * noise)."""
    rng = np.random.default_rng(RANDOM_SEED + index)
    base = rng.random((h, w, channels))
    taps = _gauss_taps(sigma)
    low = _blur_axis(_blur_axis(base, taps, 0), taps, 1)
    lo, hi = low.min(), low.max()
    low = (low - lo) / max(hi - lo, 1e-12)
    img = low + texture * (rng.random((h, w, channels)) - 0.5)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def hr_batch(n, h, w, first_index=0, **kw):
    return np.stack([hr_image(h, w, first_index + i, **kw) for i in range(n)])


def area_downsample(hr, s):
    """Exact s x s box mean (== cv2 INTER_AREA for integer factors). NHWC or HWC."""
    a = np.asarray(hr, dtype=np.float32)
    squeeze = a.ndim == 3
    if squeeze:
        a = a[None]
    n, h, w, c = a.shape
    if h % s or w % s:
        raise ValueError("HR size must be a multiple of the scale factor")
    out = a.reshape(n, h // s, s, w // s, s, c).mean(axis=(2, 4), dtype=np.float64)
    out = out.astype(np.float32)
    return out[0] if squeeze else out


def noise_batch(n, h, w, c=3, seed=RANDOM_SEED, dtype=np.float32):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype) == np.uint8:
        return rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    return rng.random((n, h, w, c), dtype=np.float32)
