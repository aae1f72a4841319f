"""The fp16 window hard-coded in csrc/metrics_mma.cu (tensor-path PSNR+SSIM): it must sum to exactly 1, stay within a few
fp16 ulp of tf.image.ssim's Gaussian, and move SSIM by far less than the 1e-4 tolerance (float64 evaluation, CPU only)."""
import os
import re
from fractions import Fraction

import numpy as np

from oracle import metrics as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200", "csrc", "metrics_mma.cu")


def _window():
    text = open(SRC).read()
    m = re.search(r"bits\[6\]\s*=\s*\{([^}]*)\}", text)
    assert m, "window bit patterns not found in metrics_mma.cu"
    bits = [int(v, 16) for v in m.group(1).replace(" ", "").split(",")]
    assert len(bits) == 6
    half = [float(np.array([b], dtype=np.uint16).view(np.float16)[0]) for b in bits]
    return np.array(half[:5] + [half[5]] + half[:5][::-1])


def _ssim64(a, b, taps):
    a, b = a.astype(np.float64), b.astype(np.float64)
    f = lambda x: om._valid_filter(x, taps)  # noqa: E731
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ma, mb = f(a), f(b)
    num0, den0 = 2 * ma * mb, ma * ma + mb * mb
    return (((num0 + c1) / (den0 + c1)) * ((2 * f(a * b) - num0 + c2) / (f(a * a + b * b) - den0 + c2))).mean(axis=(1, 2, 3))


def test_window_sums_to_one_and_tracks_the_gaussian():
    h = _window()
    assert sum(Fraction(float(v)) for v in h) == 1
    g = om.gaussian_taps(11, 1.5, np.float64)
    assert np.max(np.abs(h - g) / g) < 2.5e-3
    k2 = (np.arange(11) - 5.0) ** 2
    assert abs(float((h * k2).sum()) - float((g * k2).sum())) < 1e-4      # second moment (the window's width)


def test_window_moves_ssim_far_less_than_the_tolerance():
    h, g = _window(), om.gaussian_taps(11, 1.5, np.float64)
    rng = np.random.default_rng(0)
    n = 96
    yy, xx = np.mgrid[0:n, 0:n] / n
    smooth = np.stack([0.5 + 0.4 * np.sin(6 * xx + 3 * yy), 0.5 + 0.4 * np.cos(5 * yy), xx * yy], -1)[None]
    blocks = rng.random((1, n // 8, n // 8, 3)).repeat(8, 1).repeat(8, 2)
    for a, noise in ((smooth, 0.01), (np.full((1, n, n, 3), 0.5), 0.002), (blocks, 0.02), (rng.random((1, n, n, 3)), 0.05)):
        b = np.clip(a + noise * rng.standard_normal(a.shape), 0, 1)
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        assert np.abs(_ssim64(a32, b32, h) - _ssim64(a32, b32, g)).max() < 5e-6
